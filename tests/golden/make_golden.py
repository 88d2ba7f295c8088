"""Generate the golden vectors in tests/golden/*.npz by IMPORTING THE PYTHON REFERENCE.

Run in the build container only (the reference does not travel to the GPU box):

    rm -rf /tmp/refbuild && cp -r /root/reference /tmp/refbuild
    (cd /tmp/refbuild && python build.py)          # the reference's own cffi build
    python tests/golden/make_golden.py /tmp/refbuild

Every array below is an output of unmodified reference code (pyfilterbank.sosfiltering,
.octbank, .gammatone, .splweighting) on seeded inputs; inputs are stored beside outputs.
Environment used for the committed files: Python 3.12.3, numpy 2.3.5, scipy 1.18.1, gcc 13.3,
cffi 2.0.0 (reference extension compiled with the interpreter's CFLAGS, -O2).
"""
import os
import sys
import warnings

import numpy as np

warnings.simplefilter("ignore")
ref_root = sys.argv[1] if len(sys.argv) > 1 else "/tmp/refbuild"
sys.path.insert(0, ref_root)

import pyfilterbank.sosfiltering as sf          # noqa: E402
import pyfilterbank.octbank as ob               # noqa: E402
import pyfilterbank.butterworth as bw           # noqa: E402
import pyfilterbank.gammatone as gt             # noqa: E402
import pyfilterbank.splweighting as spl         # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def states_dict_to_array(ofb, d):
    return np.stack([np.asarray(d[f]) for f in ofb.center_frequencies])


# ---- 1. the reference's own unit-test cascade (tests/test_sosfiltering.py:11-17), seeded ----
rng = np.random.default_rng(20260101)
x = rng.standard_normal(1000)
sos = np.array([[0.5, 0.0, 0.5, 1.0, -0.8, 0.7]] * 2).astype(np.float32)
yd, sd = sf.sosfilter_double_c(x.copy(), sos)
yf, sfl = sf.sosfilter_c(x.copy(), sos)
yp, sp = sf.sosfilter_py(x.copy(), sos)
# block-wise with carried states (sosfiltering.py:143-147)
y1, s1 = sf.sosfilter_double_c(x[:333].copy(), sos)
y2, s2 = sf.sosfilter_double_c(x[333:].copy(), sos, s1)
np.savez(os.path.join(OUT, "sos_two_sections.npz"), x=x, sos=sos,
         y_double=np.array(yd), states_double=np.array(sd),
         y_float=np.array(yf), states_float=np.array(sfl),
         y_py=np.array(yp), y_blocks=np.concatenate([y1, y2]), states_blocks=np.array(s2))

# ---- 2. default fractional-octave bank (octbank.py:267-275): filter / filter_mimo_c ----
ofb = ob.FractionalOctaveFilterbank()
rng = np.random.default_rng(20260102)
xm = rng.standard_normal((1024, 2))
ym, sm = ofb.filter_mimo_c(xm)
ym2, sm2 = ofb.filter_mimo_c(xm, states=np.array(sm))          # carried states, second block
y1c, st1 = ofb.filter(xm[:, 0].copy())
y1c_b, st1b = ofb.filter(xm[:, 1].copy(), states=None)
y_ff, st_ff = ofb.filter(xm[:, 0].copy(), ffilt=True)
np.savez(os.path.join(OUT, "octbank_default.npz"), x=xm,
         sosmat=ofb.sosmat, center_frequencies=ofb.center_frequencies, band_edges=ofb.band_edges,
         effective_filter_lengths=np.array(ofb.effective_filter_lengths),
         y_mimo=np.array(ym), states_mimo=np.array(sm),
         y_mimo_block2=np.array(ym2), states_mimo_block2=np.array(sm2),
         y_filter_ch0=y1c, states_filter_ch0=states_dict_to_array(ofb, st1),
         y_filter_ch1=y1c_b,
         y_ffilt_ch0=y_ff, states_ffilt_ch0=states_dict_to_array(ofb, st_ff))

# ---- 3. design parity: other banks named by BASELINE.json configs ----
designs = {}
for name, kw in {
    "third_48k_o4": dict(sample_rate=48000, order=4, nth_oct=3.0),
    "twelfth_48k_o6": dict(sample_rate=48000, order=6, nth_oct=12.0, start_band=-67, end_band=51),
    "octave_44k_o2": dict(sample_rate=44100, order=2, nth_oct=1.0, start_band=-5, end_band=4),
}.items():
    o = ob.FractionalOctaveFilterbank(**kw)
    designs[name + "_sosmat"] = o.sosmat
    designs[name + "_cf"] = o.center_frequencies
    designs[name + "_edges"] = o.band_edges
for band, args in {"lowpass": (4, 0.1), "highpass": (4, 0.2), "bandpass": (4, 0.1, 0.2),
                   "bandstop": (2, 0.15, 0.3)}.items():
    designs["butter_" + band] = bw.butter_sos(band, *args)
np.savez(os.path.join(OUT, "designs.npz"), **designs)

# ---- 4. gammatone (gammatone.py:272-316) ----
gfb = gt.GammatoneFilterbank()
rng = np.random.default_rng(20260104)
xg = rng.standard_normal(2000)
res = list(gfb.analyze(xg))
bands = np.stack([r[0] for r in res])
gstates = np.stack([r[1] for r in res])
coef_b = np.array([c[0][0] for c in gfb._coeffs])
coef_a1 = np.array([c[1][1] for c in gfb._coeffs])
gfb4 = gt.GammatoneFilterbank(samplerate=16000, order=4, startband=-14, endband=18, density=0.5)
np.savez(os.path.join(OUT, "gammatone_default.npz"), x=xg, bands=bands, states=gstates,
         centerfrequencies=gfb.centerfrequencies, coef_b=coef_b, coef_a1=coef_a1,
         max_indices=gfb.max_indices, slopes=gfb.slopes, delay_samples=gfb.delay_samples,
         cfg4_centerfrequencies=gfb4.centerfrequencies,
         cfg4_coef_b=np.array([c[0][0] for c in gfb4._coeffs]),
         cfg4_coef_a1=np.array([c[1][1] for c in gfb4._coeffs]))

# ---- 5. SPL weighting (splweighting.py:61-80) ----
rng = np.random.default_rng(20260105)
xw = rng.standard_normal(4000)
w = {"x": xw}
for fs in (44100, 48000):
    for name, fun in (("A", spl.a_weighting_coeffs_design), ("B", spl.b_weighting_coeffs_design),
                      ("C", spl.c_weighting_coeffs_design)):
        b, a = fun(fs)
        w[f"b_{name}_{fs}"] = b
        w[f"a_{name}_{fs}"] = a
        w[f"y_{name}_{fs}"] = spl.weight_signal(xw, fs, name)
np.savez(os.path.join(OUT, "splweighting.npz"), **w)

for f in sorted(os.listdir(OUT)):
    if f.endswith(".npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))
