"""Instruction-class counts per kernel of libpfb_sos.so (what proves TMA / mbarrier / FP64 tensor-core use):

    python scripts/sass_extract.py > profiles/r02_sass_extract.txt

UTMALDG / UTMASTG = cp.async.bulk.tensor loads / stores, SYNCS = mbarrier operations, DMMA = FP64 tensor-core MMA
(m8n8k4), DFMA = FP64 FMA, VOTE = the converged warp-level barrier wait, WARPSYNC, MEMBAR.  (tcgen05 / UTCMMA has no
FP64 kind, so none is expected: the path computes in float64 / float32 recurrences.)"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyfilterbank_b200 import _lib   # noqa: E402

CLASSES = ["UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL", "FFMA", "VOTE", "WARPSYNC", "MEMBAR", "FENCE",
           "LDS", "STS", "LDG", "STG", "UTCMMA", "LDTM"]
sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
counts, name = collections.OrderedDict(), None
for line in sass.splitlines():
    if "Function :" in line:
        name = line.split("Function :")[1].strip()
        counts[name] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        counts[name]["total"] += 1
        for c in CLASSES:
            if op.startswith(c):
                counts[name][c] += 1
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("library: %s (%d kernels, %d SASS instructions)" % (os.path.basename(_lib.LIB_PATH), len(counts), tot["total"]))
print("totals : " + "  ".join("%s %d" % (c, tot[c]) for c in CLASSES))
print()
want = ("sos_tma_wide_kernel<double, double, 4", "sos_tma_kernel<double, double, 4, 1, 7", "sos_tma_kernel<double, double, 6, 1, 7",
        "sos_tma_weighted_kernel<double, double, 4, 1", "sos_tma_weighted_kernel<float, float, 4, 1", "sos_hform_kernel<double, double, 4, 11",
        "sos_carry_kernel<double, 4", "sos_tma_kernel<double, double, 4, 2, 11", "fos_tma_kernel<4", "sos_seq_kernel<double, double, 4",
        "tf_seq_kernel<double, 7", "butter_sos_kernel", "states_convert_kernel", "gt_synth")
print("%-100s %s" % ("kernel (the ones the named configs launch)", " ".join("%7s" % c[:7] for c in ["total"] + CLASSES[:13])))
for n, c in counts.items():
    d = demangle(n)
    d = re.sub(r"\(.*", "", d).replace("void ", "").replace("pfb::", "").replace("(anonymous namespace)::", "")
    if any(w in d for w in want):
        print("%-100s %s" % (d[:100], " ".join("%7d" % c[k] for k in ["total"] + CLASSES[:13])))
