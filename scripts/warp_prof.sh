#!/bin/bash
# Clocks per warp role of the output pass (PFB_WARP_PROF variant build: PFB_VARIANT=prof PFB_CFLAGS=-DPFB_WARP_PROF=1)
export PFB_WARP_PROF=1
for v in ${VARIANTS:-prof}; do
export PFB_SOS_LIB=$PWD/pyfilterbank_b200/libpfb_sos_$v.so
for nb in ${NBS:-7 4}; do
  echo "== [$v] cfg5 f64 fused nb=$nb"; PFB_TMA_NB=$nb PFB_WEIGHT_FUSED=1 python scripts/prof_fused.py cfg5 2>&1 | tail -${TAILN:-12}
done
echo "== [$v] cfg5 f32 fused"; PFB_WEIGHT_FUSED=1 python scripts/prof_fused.py cfg5f32 2>&1 | tail -9
echo "== [$v] cfg2 f64"; python scripts/prof_fused.py cfg2 2>&1 | tail -16
done
