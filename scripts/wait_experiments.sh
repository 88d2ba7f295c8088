#!/bin/bash
# Run-to-run reproducibility of the chunked gammatone path under the experimental variants of the warp-level
# barrier wait (PFB_WAIT_MODE=3 + PFB_WAIT_EXP=0..4, see mbar_wait_warp in pfb_tma.cuh).
for v in ${VARIANTS:-w3e0 w3e1 w3e2 w3e3 w3e4}; do
  echo "== variant $v"
  PFB_SOS_LIB=$PWD/pyfilterbank_b200/libpfb_sos_$v.so python scripts/determinism_probe.py gammatone 2>&1 | grep -c "True, True, True" | sed 's/^/   reproducible cases: /'
  PFB_SOS_LIB=$PWD/pyfilterbank_b200/libpfb_sos_$v.so python scripts/determinism_probe.py gammatone 2>&1 | grep -v "True, True, True" | sed 's/^/   /'
done
