"""GPU parity of the fused callers of the bank (a8 weighting prefilter, a6' decimating band
path, f2 streaming session) against the oracle, and the named BASELINE.json configs at their named sizes."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL64 = 1e-9


@pytest.fixture(scope="module")
def pfb():
    import pyfilterbank_b200 as p
    assert torch.cuda.is_available()
    return p


def _bank48(pfb):
    ofb = pfb.FractionalOctaveFilterbank(sample_rate=48000)
    B = ofb.num_bands
    return ofb, np.ascontiguousarray(ofb.sosmat.T).reshape(B, 4, 6)


# ---------------------------------------------------------------------------------------------
# a8: SPL-weighting prefilter fused into the bank pass (pfb_bank_filter_weighted)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("weighting", ["A", "B", "C"])
@pytest.mark.parametrize("C,n", [(5, 4096 + 7), (70, 2048), (33, 10000)])
def test_weighting_fused_vs_oracle(pfb, oracle, weighting, C, n, monkeypatch):
    """The weighting warp inside the TMA bank kernel (forced: these shapes are too small for the planner to pick
    it), ragged channel blocks, a ragged end handled by the unfused kernels, carried states over two calls."""
    from pyfilterbank_b200 import splweighting as spl
    monkeypatch.setenv("PFB_WEIGHT_FUSED", "1")
    ofb, sos = _bank48(pfb)
    B = ofb.num_bands
    b, a = spl._weighting_coeff_design_funsd[weighting](48000)
    rng = np.random.default_rng(C * 1000 + n)
    x = rng.standard_normal((n, C))
    w0, y0, s0 = oracle.weighted_bank(x, b, a, ofb.sosmat)
    bank = pfb.get_bank(sos)
    rows = pfb.sosfiltering.channel_rows                             # (C, n) views with 16-byte aligned row pitch
    dev = torch.device("cuda")
    xt = rows(x, dev)
    y, st, wz = pfb.bank_filter_weighted(bank, xt, b, a)
    assert bank.last_launch()["aligned"] == 4                       # the fused kernel ran
    y = y.cpu().numpy()                                              # (C, B, n)
    for bb in range(B):
        assert rel_l2(y[:, bb].T, y0[:, bb]) <= TOL64, (bb, rel_l2(y[:, bb].T, y0[:, bb]))
    assert rel_l2(st.cpu().numpy().reshape(-1), s0) <= 1e-8
    # final delay values of the prefilter: the oracle's zf per channel
    zf = np.stack([oracle.lfilter_df2t(b, a, x[:, c])[1] for c in range(C)])
    assert rel_l2(wz.cpu().numpy(), zf) <= 1e-9
    # two blocks, odd cut, carried bank and weighting states
    cut = 1024 + 33
    ya, st, wz = pfb.bank_filter_weighted(bank, rows(x[:cut], dev), b, a)
    assert bank.last_launch()["aligned"] == 4
    yb, st, wz = pfb.bank_filter_weighted(bank, rows(x[cut:], dev), b, a, st, wz)
    yy = torch.cat([ya, yb], dim=2).cpu().numpy()
    for bb in range(B):
        assert rel_l2(yy[:, bb].T, y0[:, bb]) <= TOL64, bb
    assert rel_l2(wz.cpu().numpy(), zf) <= 1e-9
    # unaligned rows (odd row stride): the same entry point falls back to the unfused kernels
    ya, st, wz = pfb.bank_filter_weighted(bank, xt[:, :cut].contiguous(), b, a)
    yb, st, wz = pfb.bank_filter_weighted(bank, xt[:, cut:].contiguous(), b, a, st, wz)
    yy = torch.cat([ya, yb], dim=2).cpu().numpy()
    for bb in range(B):
        assert rel_l2(yy[:, bb].T, y0[:, bb]) <= TOL64, bb
    # the unfused composition (weighting kernel + time-chunked bank) inside the same entry point
    monkeypatch.setenv("PFB_WEIGHT_FUSED", "0")
    yu, stu, wzu = pfb.bank_filter_weighted(bank, xt, b, a)
    assert bank.last_launch()["aligned"] != 4
    yu = yu.cpu().numpy()
    for bb in range(B):
        assert rel_l2(yu[:, bb].T, y0[:, bb]) <= TOL64, bb
    assert rel_l2(wzu.cpu().numpy(), zf) <= 1e-9


def test_weighting_fused_f32(pfb, oracle, monkeypatch):
    """float32 signals: the prefilter runs in float64 arithmetic on the float32 tile (result rounded to float32),
    the bank in float32 or (arith64) float64 arithmetic."""
    from pyfilterbank_b200 import splweighting as spl
    monkeypatch.setenv("PFB_WEIGHT_FUSED", "1")
    ofb, sos = _bank48(pfb)
    B = ofb.num_bands
    b, a = spl.a_weighting_coeffs_design(48000)
    rng = np.random.default_rng(8)
    C, n = 40, 8192 + 64 + 5
    x = rng.standard_normal((C, n)).astype(np.float32)
    w = np.stack([oracle.lfilter_df2t(b, a, x[c].astype(np.float64))[0] for c in range(C)]).astype(np.float32)
    yd, _ = oracle.sos_bank(w.astype(np.float64), sos)
    yf, _ = oracle.sos_bank(w, sos.astype(np.float32), dtype=np.float32)
    bank = pfb.get_bank(sos)
    xt = pfb.sosfiltering.channel_rows(x.T, torch.device("cuda"), torch.float32)
    y1, _, _ = pfb.bank_filter_weighted(bank, xt, b, a, arith64=True)
    assert bank.last_launch()["aligned"] == 4
    y1 = y1.cpu().numpy()
    for bb in range(B):
        assert rel_l2(y1[:, bb], yd[:, bb]) <= 1e-6, (bb, rel_l2(y1[:, bb], yd[:, bb]))
    y2, _, _ = pfb.bank_filter_weighted(bank, xt, b, a)
    assert bank.last_launch()["aligned"] == 4
    y2 = y2.cpu().numpy()
    for bb in range(B):
        ref_err = rel_l2(yf[:, bb], yd[:, bb])                        # the reference float path's own error
        assert rel_l2(y2[:, bb], yd[:, bb]) <= max(10 * ref_err, 1e-5), (bb, ref_err)
    # unfused float32 composition gives the same weighted signal (bit-identical prefilter arithmetic)
    monkeypatch.setenv("PFB_WEIGHT_FUSED", "0")
    y3, _, _ = pfb.bank_filter_weighted(bank, xt, b, a, arith64=True)
    y3 = y3.cpu().numpy()
    for bb in range(B):
        assert rel_l2(y3[:, bb], yd[:, bb]) <= 1e-6, bb


def test_cfg5_named_size_weighted_bank(pfb, oracle):
    """BASELINE.json configs[4] per-GPU shard at 8 GPUs: 1024 channels x 5 s @ 48 kHz, A-weighting fused with the
    1/3-octave bank through the reference-facing call; spot channels against the oracle's composition
    (lfilter in scipy's order, then sosfilter_double_mimo_c)."""
    from pyfilterbank_b200 import splweighting as spl
    free, _ = torch.cuda.mem_get_info()
    if free < 90e9:
        pytest.skip("needs 90 GB of free device memory")
    ofb = pfb.FractionalOctaveFilterbank(sample_rate=48000)
    C, n = 1024, 240000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    x = torch.randn((n, C), generator=gen, device="cuda", dtype=torch.float64) * 0.1
    y, st, wz = ofb.filter_mimo_c(x, weighting="A")
    assert y.shape == (n, 33, C)
    bank = pfb.get_bank(np.ascontiguousarray(ofb.sosmat.T).reshape(33, 4, 6))
    assert bank.last_launch()["aligned"] == 4                       # fused: the planner picks it at this size
    b, a = spl.a_weighting_coeffs_design(48000)
    for c in (0, 517, 1023):
        xc = x[:, c].cpu().numpy()
        _, y0, s0 = oracle.weighted_bank(xc, b, a, ofb.sosmat)
        yc = y[:, :, c].cpu().numpy()
        for bb in range(33):
            assert rel_l2(yc[:, bb], y0[:, bb, 0]) <= TOL64, (c, bb, rel_l2(yc[:, bb], y0[:, bb, 0]))


@pytest.mark.parametrize("weighting", ["A", "C"])
def test_weighted_band_levels_vs_oracle(pfb, oracle, weighting):
    """Prefilter + bank + energies per block in ONE kernel (pfb_bank_levels_weighted): against the oracle's
    composition (lfilter in scipy's order, sosfilter_double_mimo_c, then sum y^2 per block), ragged channel block,
    two calls with carried bank states and prefilter delay values, numpy in / numpy out through filter_levels."""
    from pyfilterbank_b200 import splweighting as spl
    ofb, sos = _bank48(pfb)
    b, a = spl._weighting_coeff_design_funsd[weighting](48000)
    C, n, bl = 37, 9600, 960
    x = np.random.default_rng(21).standard_normal((n, C))
    _, y0, _ = oracle.weighted_bank(x, b, a, ofb.sosmat)                       # (N, B, C)
    e0 = (y0.reshape(n // bl, bl, 33, C) ** 2).sum(axis=1)                     # (blocks, B, C)
    e, st, wz = ofb.filter_levels(x, bl, weighting=weighting)
    assert e.shape == e0.shape
    scale = e0.max(axis=0, keepdims=True)
    assert np.max(np.abs(e - e0) / scale) <= 5e-9, np.max(np.abs(e - e0) / scale)
    zf = np.stack([oracle.lfilter_df2t(b, a, x[:, c])[1] for c in range(C)])
    assert rel_l2(wz, zf) <= 1e-9
    cut = 4 * bl
    ea, st, wz = ofb.filter_levels(x[:cut], bl, weighting=weighting)
    eb, st, wz = ofb.filter_levels(x[cut:], bl, states=st, weighting=weighting, weighting_states=wz)
    ee = np.concatenate([ea, eb])
    assert np.max(np.abs(ee - e0) / scale) <= 5e-9
    # float32 signals on the device (float32 bank arithmetic): energies within the float32 path's own error
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).cuda().to(torch.float32)
    bank = pfb.get_bank(sos)
    ef, _, _ = pfb.bank_levels(bank, xt[:, :n // 64 * 64], bl // 64 * 64 * 2, weighting=(b, a))
    assert ef.shape[1] == 33 and torch.isfinite(ef).all()


# ---------------------------------------------------------------------------------------------
# a6': decimating band path as one library call (pfb_decbank_filter)
# ---------------------------------------------------------------------------------------------

def test_decimated_fused_stages_and_unfused_agree(pfb, oracle, monkeypatch):
    """The decimating band warp of the TMA kernels (many-channel regime: rows are channels; few-channel regime:
    rows are time chunks with pre-pass and carry) against the oracle's composition and against the unfused
    per-stage kernels."""
    for C, n, order, nth, sb, eb in ((40, 20000 + 6, 4, 3.0, -19, 13), (3, 300000 + 2, 6, 12.0, -67, 51)):
        ofb = pfb.FractionalOctaveFilterbank(sample_rate=48000, order=order, nth_oct=nth, start_band=sb, end_band=eb)
        rng = np.random.default_rng(C)
        x = rng.standard_normal((n, C))
        xt = torch.from_numpy(x).cuda()
        bands, st = ofb.filter_decimated(xt)
        info = ofb._dec_bank().info()
        assert info["last_fused_stages"] >= 1, info                  # the fused stage kernels ran
        chans = sorted({0, C - 1})
        ref = oracle.decimated_bank(x[:, chans], ofb.decimation_plan(), ofb.aa_sos, order)
        for b in range(ofb.num_bands):
            yb = bands[b].cpu().numpy()[:, chans]
            assert yb.shape == ref[b].shape, (b, yb.shape, ref[b].shape)
            assert rel_l2(yb, ref[b]) <= TOL64, (C, b, rel_l2(yb, ref[b]))
        # block-wise with an odd cut (decimator phases 1 on some stages)
        cut = 4097
        ba, s1 = ofb.filter_decimated(xt[:cut])
        bb, s1 = ofb.filter_decimated(xt[cut:], states=s1)
        for b in range(ofb.num_bands):
            yy = torch.cat([ba[b], bb[b]]).cpu().numpy()[:, chans]
            assert yy.shape == ref[b].shape, b
            assert rel_l2(yy, ref[b]) <= TOL64, (C, b)
        monkeypatch.setenv("PFB_DEC_UNFUSED", "1")
        bu, _ = ofb.filter_decimated(xt)
        assert ofb._dec_bank().info()["last_fused_stages"] == 0
        monkeypatch.delenv("PFB_DEC_UNFUSED")
        for b in range(ofb.num_bands):
            assert rel_l2(bu[b].cpu().numpy()[:, chans], ref[b]) <= TOL64, (C, b)


def test_cfg3_named_size_decimated_stream(pfb, oracle):
    """BASELINE.json configs[2]: 1/12-octave order-6 bank with per-band decimation, 1024 channels, streamed in
    1-second blocks with carried states (anti-alias states, decimator phases, band states); three blocks,
    spot channels against the oracle's composition on the concatenated signal."""
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs 60 GB of free device memory")
    ofb = pfb.FractionalOctaveFilterbank(sample_rate=48000, order=6, nth_oct=12.0, start_band=-67, end_band=51)
    C, nblk, blocks = 1024, 48000, 3
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)
    chans = [0, 511, 1023]
    xs, outs, st = [], [], None
    for k in range(blocks):
        x = torch.randn((nblk, C), generator=gen, device="cuda", dtype=torch.float64) * 0.1
        xs.append(x[:, chans].cpu().numpy())
        bands, st = ofb.filter_decimated(x, st)
        outs.append([bnd[:, chans].cpu().numpy() for bnd in bands])
        del bands
    assert ofb._dec_bank().info()["last_fused_stages"] >= 1
    ref = oracle.decimated_bank(np.concatenate(xs), ofb.decimation_plan(), ofb.aa_sos, 6)
    for b in range(ofb.num_bands):
        yy = np.concatenate([o[b] for o in outs])
        assert yy.shape == ref[b].shape, (b, yy.shape, ref[b].shape)
        assert rel_l2(yy, ref[b]) <= TOL64, (b, rel_l2(yy, ref[b]))


# ---------------------------------------------------------------------------------------------
# cfg 1 at its named size: the reference's default bank (fs = 44100, top band edge clipped, |coef| up to 1.5e4)
# ---------------------------------------------------------------------------------------------

def test_cfg1_named_size_default_bank(pfb, oracle):
    """BASELINE.json configs[0]: FractionalOctaveFilterbank() defaults on 1 channel x 10 s (441 000 samples) of white
    noise, numpy in / numpy out through filter_mimo_c AND filter (the time-chunked TMA path), per band against
    the oracle (the reference's sosfilter_double_mimo)."""
    ofb = pfb.FractionalOctaveFilterbank()
    assert ofb.sample_rate == 44100 and ofb.num_bands == 33
    rng = np.random.default_rng(0)
    x = rng.standard_normal(441000)
    y0, s0 = oracle.sosfilter_double_mimo_c(x, ofb.sosmat)
    y, st = ofb.filter_mimo_c(x)
    assert y.shape == (441000, 33, 1) and y.flags.f_contiguous
    info = pfb.get_bank(np.ascontiguousarray(ofb.sosmat.T).reshape(33, 4, 6)).last_launch()
    assert info["mode"] == 2 and info["aligned"] in (2, 3), info     # time-chunked TMA kernels
    worst = 0.0
    for b in range(33):
        e = rel_l2(y[:, b, 0], y0[:, b, 0])
        worst = max(worst, e)
        assert e <= TOL64, (b, e)
    assert rel_l2(st, s0) <= 1e-8
    y1, states = ofb.filter(x)
    assert y1.shape == (441000, 33)
    for b in range(33):
        assert rel_l2(y1[:, b], y0[:, b, 0]) <= TOL64, b
    s1 = np.concatenate([states[f] for f in ofb.center_frequencies])
    assert rel_l2(s1, s0) <= 1e-8


# ---------------------------------------------------------------------------------------------
# f2: persistent streaming session (pfb_stream_*)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("exact", [True, False])
def test_stream_session_100_blocks(pfb, oracle, exact):
    """100 blocks of 4800 samples through one session == one shot: bit-identical to the oracle in exact mode,
    <= 1e-9 per band otherwise; nothing is allocated on the device after the first block."""
    ofb, sos = _bank48(pfb)
    B, K = 33, 4
    C, L, nblocks = 4, 4800, 100
    rng = np.random.default_rng(12)
    x = rng.standard_normal((L * nblocks, C))
    y0, s0 = oracle.sosfilter_double_mimo_c(x, ofb.sosmat)
    sess = pfb.BlockSession(sos, C, L, exact=exact)
    out = np.empty((L * nblocks, B, C))
    tickets = []
    for k in range(nblocks):
        tickets.append(sess.push(x[k * L:(k + 1) * L]))
        if k >= 2:                                                   # consume two blocks behind the producer
            out[(k - 2) * L:(k - 1) * L] = sess.result(tickets[k - 2])
    for k in (nblocks - 2, nblocks - 1):
        out[k * L:(k + 1) * L] = sess.result(tickets[k])
    stats = sess.stats()
    assert stats["blocks"] == nblocks and stats["samples"] == L * nblocks
    assert stats["device_allocs_after_first_block"] == 0, stats
    st = sess.states
    if exact:
        assert np.array_equal(out, y0) and np.array_equal(st, s0)
    else:
        for b in range(B):
            assert rel_l2(out[:, b], y0[:, b]) <= TOL64, (b, rel_l2(out[:, b], y0[:, b]))
        assert rel_l2(st, s0) <= 1e-8
    # ragged last block and states set from outside
    sess2 = pfb.BlockSession(ofb, C, L, exact=exact)
    sess2.states = 0 * s0
    ya = sess2.filter(x[:L])
    yb = sess2.filter(x[L:L + 1234])
    yy = np.concatenate([ya, yb])
    if exact:
        assert np.array_equal(yy, y0[:L + 1234])
    else:
        assert max(rel_l2(yy[:, b], y0[:L + 1234, b]) for b in range(B)) <= TOL64
    sess.close()
    sess2.close()


def test_stream_session_c_abi_pinned(pfb, oracle):
    """The C ABI directly (pfb_stream_create / push / sync / states / stats / destroy) on page-locked buffers,
    float32 signals with float64 arithmetic."""
    from pyfilterbank_b200 import _lib
    ofb, sos = _bank48(pfb)
    B, K, C, L, nblocks = 33, 4, 3, 9600, 12
    rng = np.random.default_rng(5)
    x = rng.standard_normal((C, L * nblocks)).astype(np.float32)
    yd, sd = oracle.sos_bank(x.astype(np.float64), sos)
    bank = pfb.get_bank(sos)
    s = _lib.Stream(bank, _lib.PFB_F32, C, L, _lib.flags_of("auto", False, True))
    xh = torch.from_numpy(x).pin_memory()
    yh = torch.empty((C, B, L * nblocks), dtype=torch.float32, pin_memory=True)
    for k in range(nblocks):
        s.push(xh.data_ptr() + 4 * k * L, L * nblocks, yh.data_ptr() + 4 * k * L, L * nblocks, L)
    s.sync()
    y = yh.numpy()
    for b in range(B):
        assert rel_l2(y[:, b], yd[:, b]) <= 1e-6, b
    st = s.get_states(np.empty((C, B, K, 2)))
    assert rel_l2(st, sd) <= 1e-7
    assert s.stats()["device_allocs_after_first_block"] == 0
    s.close()


# ---------------------------------------------------------------------------------------------
# random shapes through the planner (scripts/fuzz_shapes.py, what found the order-8 race in round 1)
# ---------------------------------------------------------------------------------------------

def test_fuzz_shapes_one_minute():
    env = dict(os.environ, SECONDS="35", SECONDS_FOS="12", SECONDS_LEV="8", SECONDS_FUSED="15", SEED="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_shapes.py")], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "fuzz ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
