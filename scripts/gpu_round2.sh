#!/bin/bash
# One GPU-box visit of round 2.  Sections via $RUN (space separated): fused tests configs bench ncu
TAG=${1:-r02a}
RUN=${RUN:-"fused tests configs bench"}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; free -g >> $O/${TAG}_gpu.txt
has() { [[ " $RUN " == *" $1 "* ]]; }
if has fused; then
  timeout 1500 python -m pytest tests/test_gpu_fused.py -q -x --deselect tests/test_gpu_fused.py::test_fuzz_shapes_one_minute ${PYTEST_ARGS} > $O/${TAG}_pytest_fused.log 2>&1; echo "pytest fused rc=$?" | tee -a $O/${TAG}_pytest_fused.log
  tail -25 $O/${TAG}_pytest_fused.log
fi
if has tests; then
  timeout 2400 python -m pytest tests -m gpu -q ${PYTEST_ALL_ARGS:--x} > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
  tail -15 $O/${TAG}_pytest.log
fi
if has configs; then
  timeout 900 python scripts/bench_configs.py ${CONFIG_CASES:-cfg1 cfg3_dec cfg3_unfused cfg3_stream cfg4 cfg5 cfg5_f32} > $O/${TAG}_configs.jsonl 2> $O/${TAG}_configs.err; echo "configs rc=$?"
  cut -c1-400 $O/${TAG}_configs.jsonl; tail -5 $O/${TAG}_configs.err
fi
if has sweep; then
  timeout 900 python scripts/bench_configs.py cfg5_sweep > $O/${TAG}_cfg5_sweep.jsonl 2> $O/${TAG}_cfg5_sweep.err; echo "sweep rc=$?"
  cut -c1-330 $O/${TAG}_cfg5_sweep.jsonl; tail -5 $O/${TAG}_cfg5_sweep.err
fi
if has bench; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"
  timeout 900 python bench.py ${BENCH_ARGS} > $O/${TAG}_bench_f64.json 2> $O/${TAG}_bench_f64.err; echo "bench f64 rc=$?"
  tail -3 $O/${TAG}_bench_f64.err
  cut -c1-6000 $O/${TAG}_bench_f64.json
fi
if has bench32; then
  timeout 900 python bench.py --dtype f32 --no-cpu --no-configs > $O/${TAG}_bench_f32.json 2> $O/${TAG}_bench_f32.err; echo "bench f32 rc=$?"
  cut -c1-3000 $O/${TAG}_bench_f32.json
fi
if has ncu; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --no-f32 > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu list rc=$?"
  mkdir -p /tmp/ncu
  for case in ${NCU_CASES:-cfg5 cfg3}; do
    # the reports stay on the box (gpurun_out/ is limited to 64 MiB): only their text summaries come back
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-sos_tma|sos_hform|tf_seq}" -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-6} -f \
        -o /tmp/ncu/${TAG}_full_${case} python scripts/prof_fused.py $case > $O/${TAG}_ncu_${case}.log 2>&1; echo "ncu $case rc=$?"
    python scripts/ncu_summary.py /tmp/ncu/${TAG}_full_${case}.ncu-rep > $O/${TAG}_full_${case}.txt 2>&1
    python scripts/ncu_hot.py /tmp/ncu/${TAG}_full_${case}.ncu-rep ${NCU_TOP:-45} > $O/${TAG}_hot_${case}.txt 2>&1
    sz=$(stat -c %s /tmp/ncu/${TAG}_full_${case}.ncu-rep 2>/dev/null || echo 0)
    if [ "$sz" -lt 12000000 ] && [ -n "$NCU_KEEP" ]; then cp /tmp/ncu/${TAG}_full_${case}.ncu-rep $O/; fi
  done
fi
