"""Golden vectors for SURVEY.md section 8 row f4 (coefficient design on the device, 'py' state convention), made by
IMPORTING THE PYTHON REFERENCE (see make_golden.py for the build recipe of /tmp/refbuild):

    python tests/golden/make_golden_f4.py /tmp/refbuild     ->  tests/golden/f4_design_and_py_states.npz
"""
import os
import sys
import warnings

import numpy as np

warnings.simplefilter("ignore")
sys.path.insert(0, sys.argv[1] if len(sys.argv) > 1 else "/tmp/refbuild")
import pyfilterbank.sosfiltering as sf          # noqa: E402
import pyfilterbank.octbank as ob               # noqa: E402
import pyfilterbank.butterworth as bw           # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
d = {}
# 1. butter_sos for every band kind (butterworth.py:16-58), a few orders and frequency pairs
cases = []
for kind in ("lowpass", "highpass", "bandpass", "bandstop"):
    for L in (2, 4, 8):
        for v1, v2 in ((0.01, 0.02), (0.1, 0.2), (0.2, 0.45), (0.0003, 0.00035)):
            sos = bw.butter_sos(kind, L, v1, v2 if kind in ("bandpass", "bandstop") else 0.0)
            cases.append((kind, L, v1, v2))
            d["design_%s_%d_%g_%g" % (kind, L, v1, v2)] = sos
d["design_cases"] = np.array(["%s %d %r %r" % c for c in cases])
# 2. whole banks (octbank.py:132-173): default 44.1 kHz bank (top band clipped), 1/12-octave order 6 at 48 kHz
d["bank_default"] = ob.FractionalOctaveFilterbank().sosmat
d["bank_twelfth_48k_o6"] = ob.FractionalOctaveFilterbank(sample_rate=48000, order=6, nth_oct=12.0, start_band=-67, end_band=51).sosmat
# 3. sosfilter_py (sosfiltering.py:330-379): one shot and two blocks with carried dict states
rng = np.random.default_rng(44)
x = rng.standard_normal(6000)
ofb = ob.FractionalOctaveFilterbank(sample_rate=48000)
for name, sos in (("band5", ofb.sosmat[:, 5].reshape(4, 6)), ("band30", ofb.sosmat[:, 30].reshape(4, 6)),
                  ("unnormalised", np.array([[0.5, 0.1, 0.3, 2.0, -1.2, 0.9], [1.0, -0.4, 0.2, 0.5, 0.1, 0.2]]))):
    y, st = sf.sosfilter_py(x.copy(), sos)
    ya, sta = sf.sosfilter_py(x[:2501].copy(), sos)
    mid = np.stack([np.array(sta[i]) for i in range(len(sos))])     # (sosfilter_py updates the dict it is given)
    yb, stb = sf.sosfilter_py(x[2501:].copy(), sos, sta)
    assert np.allclose(np.concatenate([ya, yb]), y, rtol=0, atol=1e-12)
    d["py_%s_sos" % name] = sos
    d["py_%s_y" % name] = y
    d["py_%s_states" % name] = np.stack([st[i] for i in range(len(sos))])
    d["py_%s_states_mid" % name] = mid
d["py_x"] = x
# 4. FractionalOctaveFilterbank.filter with filterfun 'py' (octbank.py:402-404, 436-476): dict of dict states
ofb.set_filterfun('py')
y, st = ofb.filter(x[:1200])
d["py_bank_y"] = y
d["py_bank_states"] = np.stack([np.stack([st[f][i] for i in range(4)]) for f in ofb.center_frequencies])
np.savez_compressed(os.path.join(OUT, "f4_design_and_py_states.npz"), **d)
print("wrote", len(d), "arrays")
