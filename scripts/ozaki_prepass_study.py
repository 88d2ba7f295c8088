"""CPU study for DESIGN.md section 8 item 1: how many 7-bit integer slices would an exact-product (Ozaki-style)
INT8 tensor-core form of the pre-pass GEMM z = H x need to stay under the 1e-11 budget of the float64 path?
H = chunk impulse response of one band's state vector (2K x W), x = white noise.  numpy only, no device."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyfilterbank_b200.octbank import FractionalOctaveFilterbank

def state_impulse_response(sos, W):
    """H[:, t] = state vector (w1, w2 per section) W - t samples after a unit input sample (DF-II cascade)."""
    K = sos.shape[0]
    st = np.zeros((K, 2), np.longdouble)
    H = np.zeros((2 * K, W), np.longdouble)
    x = np.longdouble(1.0)
    for t in range(W):
        v = x
        for k in range(K):
            b0, b1, b2, _, a1, a2 = [np.longdouble(c) for c in sos[k]]
            w0 = v - a1 * st[k, 0] - a2 * st[k, 1]
            v = b0 * w0 + b1 * st[k, 0] + b2 * st[k, 1]
            st[k, 1], st[k, 0] = st[k, 0], w0
        x = np.longdouble(0.0)
        H[:, W - 1 - t] = st.reshape(-1)
    return H

def slices(v, nslices, bits=7):
    """v (float64 vector) -> exponent e and integer slices with v ~ 2^e * sum_i s_i 2^(-bits (i + 1)), |s_i| <= 2^(bits-1)."""
    e = int(np.ceil(np.log2(np.max(np.abs(v)) + 1e-300))) + 1
    r = v / 2.0 ** e
    out = []
    for i in range(nslices):
        r = r * 2.0 ** bits
        s = np.rint(r)
        out.append(s.astype(np.int64))
        r = r - s
    return e, out

def main():
    try:
        ofb = FractionalOctaveFilterbank(sample_rate=48000)
        sosmat = np.asarray(ofb.sosmat)
    except Exception as exc:              # constructor needs no device, but be explicit if that ever changes
        print("bank design unavailable:", exc); return
    K = sosmat.shape[0] // 6
    rng = np.random.default_rng(0)
    W = 8192
    x = rng.standard_normal(W)
    for band in (0, 10, 20, sosmat.shape[1] - 1):
        sos = sosmat[:, band].reshape(K, 6)
        H = state_impulse_response(sos, W)
        z_ref = (H @ x.astype(np.longdouble))
        z64 = (H.astype(np.float64) @ x)
        line = "band %2d: float64 dot %.1e" % (band, float(np.max(np.abs(z64 - z_ref) / np.max(np.abs(z_ref)))))
        for ns in (4, 5, 6, 7, 8):
            ex, xs = slices(x, ns)
            z = np.zeros(2 * K, np.longdouble)
            nprod = 0
            for r in range(2 * K):
                eh, hs = slices(H[r].astype(np.float64), ns)
                acc = np.longdouble(0)
                for i in range(ns):
                    for j in range(ns - i):            # triangular: drop products below 2^(-7 (ns + 1))
                        acc += np.longdouble(int(np.dot(hs[i], xs[j]))) * np.longdouble(2.0) ** (-7 * (i + j + 2))
                        nprod += (r == 0)
                z[r] = acc * np.longdouble(2.0) ** (eh + ex)
            line += " | %d slices (%d products) %.1e" % (ns, nprod, float(np.max(np.abs(z - z_ref) / np.max(np.abs(z_ref)))))
        print(line, flush=True)

if __name__ == "__main__":
    main()
