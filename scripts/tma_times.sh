#!/bin/bash
# per-kernel device times of the cfg-2 workload (ncu launch list, cold-cache, serialised)
for dt in ${DTS:-f64}; do
DT=$dt N=2880000 IT=2 python scripts/prof_case.py >/dev/null 2>&1 && DT=$dt N=2880000 IT=2 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sos_|carry" --csv --log-file gpurun_out/tma_launches.csv python scripts/prof_case.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/tma_launches.csv')) if len(r)>10][1:]
n=len(rows)//2
for r in rows[n:]:
    print("   %-70s grid %-14s %8.3f ms" % (r[4][:70], r[8], float(r[-1])/1e6))
print("   total %.3f ms" % (sum(float(r[-1]) for r in rows[n:])/1e6))
PY
done
