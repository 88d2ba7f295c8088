# knobs of the split schedule of the decimating band path (pfb_decbank::split_filter) on cfg 3 (1024 ch x 4 s)
for j0 in 0 2 4 8; do for share in 0.5 0.125; do
echo "stage0 J=$j0 share=$share"; PFB_DEC_SPLIT_TRACE=1 PFB_DEC_S0_J=$j0 PFB_DEC_SPLIT_SHARE=$share python scripts/bench_configs.py cfg3_dec 2>&1 | grep 'dec split\|"case"' | tail -2 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ms %.2f frac %.3f kernels %d' % (d['ms'], d['frac_of_hbm_peak'], d['dec']['last_kernels']))
    else: print('  '+l.strip()[:420])"
done; done
