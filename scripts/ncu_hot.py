"""Hottest SASS instructions (warp-stall samples) of every kernel in an .ncu-rep, with the dominant stall reasons:
a text digest small enough to bring back from the GPU box (gpurun_out/ is capped at 64 MiB).

    python scripts/ncu_hot.py report.ncu-rep [top_n] [kernel-substring] > profiles/xxx_hot.txt
"""
import csv
import io
import subprocess
import sys


def main(path, top=40, want=""):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    i = 0
    seen = set()
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            hdr = rows[i + 1]
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if len(rows[j]) == len(hdr):
                    body.append(rows[j])
                j += 1
            i = j
            if want not in name or name in seen:
                continue
            seen.add(name)
            col = {h: k for k, h in enumerate(hdr)}
            samp = col["# Samples"]
            stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            total = sum(int(r[samp] or 0) for r in body)
            print("\n== %s\n   %d instructions, %d stall samples" % (name[:160], len(body), total))
            agg = {h: sum(int(r[col[h]] or 0) for r in body) for h in stalls}
            print("   by reason: " + ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(total, 1))
                                               for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
            order = sorted(range(len(body)), key=lambda k: -int(body[k][samp] or 0))[:top]
            for k in sorted(order):
                r = body[k]
                n = int(r[samp] or 0)
                why = sorted(((int(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
                print("   #%5d %5.1f%%  %-70s %s" % (k, 100.0 * n / max(total, 1), r[col["Source"]].strip()[:70],
                                                     " ".join("%s:%d" % (h, v) for v, h in why if v)))
        else:
            i += 1


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else "")
