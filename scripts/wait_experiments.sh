#!/bin/bash
# Run-to-run reproducibility of the chunked gammatone path under experimental variants of the warp-level barrier wait.
# HISTORICAL: the PFB_WAIT_EXP switches (delay, proxy fence, repeated try_waits, release by every lane, old wait in one
# of the two passes) lived in pfb_tma.cuh / pfb_fos_tma.cuh only for commit 1fb01aa..; results: profiles/r02p*_wait_experiments.txt.
# Variant libraries are built with  PFB_VARIANT=name PFB_CFLAGS="-D..." python -m pyfilterbank_b200._build .
for v in ${VARIANTS:-w3e0 w3e1 w3e2 w3e3 w3e4}; do
  echo "== variant $v"
  PFB_SOS_LIB=$PWD/pyfilterbank_b200/libpfb_sos_$v.so python scripts/determinism_probe.py gammatone 2>&1 | grep -c "True, True, True" | sed 's/^/   reproducible cases: /'
  PFB_SOS_LIB=$PWD/pyfilterbank_b200/libpfb_sos_$v.so python scripts/determinism_probe.py gammatone 2>&1 | grep -v "True, True, True" | sed 's/^/   /'
done
