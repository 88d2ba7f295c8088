#!/bin/bash
# One GPU-box visit: parity tests, bench lines (f64 / f32 / reference arm), ncu launch list of the
# bench command and one full ncu capture of the two big kernels.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
  tail -5 $O/${TAG}_pytest.log
fi
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/${TAG}_bench_f64.json 2> $O/${TAG}_bench_f64.err; echo "bench f64 rc=$?"
timeout 900 python bench.py --dtype f32 --no-cpu > $O/${TAG}_bench_f32.json 2> $O/${TAG}_bench_f32.err; echo "bench f32 rc=$?"
cat $O/${TAG}_bench_f64.json $O/${TAG}_bench_f32.json $O/${TAG}_bench_ref.json | cut -c1-1800
if [ -z "$SKIP_NCU" ]; then
  # launch list of the bench command itself (cold-cache, serialised: shares, not absolutes)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu list rc=$?"
  # full capture of the pre-pass and output kernels on a quarter-length signal (same grid shape per chunk group)
  DT=f64 C=64 N=${NCU_N:-1440000} IT=2 timeout 1200 ncu --set full --clock-control none --import-source on \
      -k regex:"sos_tma|sos_carry|sos_hform" -s 3 -c 3 -f -o $O/${TAG}_full_f64 python scripts/prof_case.py > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
  DT=f32 C=64 N=${NCU_N:-1440000} IT=2 timeout 1200 ncu --set full --clock-control none --import-source on \
      -k regex:"sos_tma|sos_hform" -s 2 -c 2 -f -o $O/${TAG}_full_f32 python scripts/prof_case.py > $O/${TAG}_ncu_full32.log 2>&1; echo "ncu full32 rc=$?"
fi
if [ -n "$RUN_UBENCH" ]; then
  for b in $RUN_UBENCH; do timeout 120 scripts/ubench/$b.bin > $O/${TAG}_ubench_$b.txt 2>&1; echo "ubench $b rc=$?"; cat $O/${TAG}_ubench_$b.txt | tail -30; done
fi
