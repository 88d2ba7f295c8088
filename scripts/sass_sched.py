"""Static issue-cycle estimate of a SASS range from the control words (stall counts) ptxas encoded.

    python scripts/sass_sched.py <file.sass | lib.so> <kernel-substring> [addr_lo addr_hi]

Per instruction: stall count (bits 41-44 of the high word), yield, write/read barrier ids, wait mask.  The sum of
stall counts over a straight-line range is the issue time of one warp running alone when no scoreboard wait
blocks (fixed-latency pipes).  Used to compare loop schedules without a GPU.
"""
import re
import subprocess
import sys


def dump(path, kern):
    if path.endswith(".sass"):
        return open(path).read()
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    return out


def parse(text, kern):
    funcs = re.split(r"\n\s*Function : ", text)
    res = []
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if kern not in name:
            continue
        lines = f.split("\n")
        ins = []
        i = 0
        while i < len(lines):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", lines[i])
            if m and i + 1 < len(lines):
                m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
                if m2:
                    hi = int(m2.group(1), 16)
                    ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xf, (hi >> 45) & 1,
                                (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f))
                    i += 2
                    continue
            i += 1
        res.append((name, ins))
    return res


if __name__ == "__main__":
    path, kern = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
    verbose = "-v" in sys.argv
    for name, ins in parse(dump(path, kern), kern):
        sel = [x for x in ins if lo <= x[0] <= hi]
        tot = sum(x[2] for x in sel)
        fp = sum(1 for x in sel if re.match(r"(@!?U?P\d+ )?(DFMA|DMUL|DADD|DMMA|FFMA|FMUL|FADD)", x[1]))
        print("%s\n  instructions %d, FP %d, sum of stall counts %d" % (name[:150], len(sel), fp, tot))
        if verbose:
            for a, t, st, y, wb, rb, wm in sel:
                print("  %05x st=%2d y=%d wb=%d rb=%d wait=%02x  %s" % (a, st, y, wb, rb, wm, t))
