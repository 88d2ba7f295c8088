"""Build libpfb_sos.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m pyfilterbank_b200._build [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# experiment builds: PFB_VARIANT=name PFB_CFLAGS="-DX=1 ..." -> libpfb_sos_name.so (loaded with PFB_SOS_LIB=<path>)
VARIANT = os.environ.get("PFB_VARIANT", "")
OBJ = os.path.join(CSRC, "build" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "libpfb_sos%s.so" % ("_" + VARIANT if VARIANT else ""))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v"] + os.environ.get("PFB_CFLAGS", "").split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inl"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "pfb_sos.h"))
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ, src[:-3] + ".o")
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj[:-2] + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        return src

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for done in ex.map(compile_one, jobs):
            if verbose:
                print("compiled", done)
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in sources()]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr)
        if verbose:
            print("linked", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
