"""Registers / spills per kernel from the ptxas logs of the last build (pyfilterbank_b200/csrc/build/*.ptxas.log).

    python scripts/ptxas_summary.py [substring ...]
"""
import glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyfilterbank_b200", "csrc", "build")
rows = []
for f in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    t = open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", t, re.S):
        rows.append((m.group(1), int(m.group(5)), int(m.group(3)), int(m.group(4))))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.split("\n")
for (mn, regs, ss, sl), dn in zip(rows, names):
    short = re.sub(r"\(.*", "", dn).replace("void pfb::", "")
    if all(s in short for s in sys.argv[1:]):
        print("%-80s regs %3d  spill st/ld %d/%d" % (short, regs, ss, sl))
