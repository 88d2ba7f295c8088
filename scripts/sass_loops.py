"""List the loops (backward branches) of one kernel in a cubin/.o and their instruction mix.
usage: python scripts/sass_loops.py <obj> <mangled-name-substring>"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", blk):
        ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name, "instructions:", len(ins))
    for a, t in ins:
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            body = [x for x in ins if lo <= x[0] <= a]
            mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
            pred = sum(1 for x in body if x[1].startswith("@"))
            print("  loop 0x%x..0x%x: %d instr (%d predicated): %s" % (lo, a, len(body), pred,
                  ", ".join("%s %d" % kv for kv in mix.most_common(14))))
