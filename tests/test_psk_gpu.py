"""GPU parity of the single- and multi-carrier DPSK soft demodulators (csrc/psk_demod.cu, through the C ABI) against
the oracle on the same received samples: LLR words bit-identical (sequential fp32 correlations are reproduced in
order; libm calls are restated), for every modulation, symbol rate, reference mode, CFO / phase compensation,
ragged lengths, silence, host and device memory, and full-size frames of SURVEY config 4 with LDPC decoding."""
import numpy as np
import pytest

import oracleapi as O
import refapi as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from projectultra_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def same_bits(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return a.shape == b.shape and bool((a == b).all())


def noisy(x, snr_db, rng):
    p = float(np.mean(x.astype(np.float64) ** 2))
    return (x + rng.standard_normal(len(x)).astype(np.float32) * np.float32(np.sqrt(p / 10 ** (snr_db / 10)))).astype(np.float32)


@pytest.mark.parametrize("mod", [0, 1, 2])
@pytest.mark.parametrize("sps", [384, 192, 100])
def test_sc_dpsk_llrs_bit_identical(ctx, mod, sps):
    from projectultra_b200 import capi
    rng = np.random.default_rng(31 * mod + sps)
    cfg = capi.dpsk_config(mod, sps)
    dem = capi.DpskDemodulator(ctx, cfg)
    data = rng.integers(0, 256, 30, dtype=np.uint8)
    tx = capi.dpsk_tx(cfg, data, 0)
    tx = (tx * np.float32(0.5 / np.abs(tx).max())).astype(np.float32)
    frames = np.stack([noisy(tx, snr, rng) for snr in (25.0, 8.0, 0.0, -6.0, -11.0)] + [np.zeros_like(tx), tx * np.float32(1e-4)])
    B, L = frames.shape
    start = 39 * sps
    for ds, ref_mode, use_comp in ((start, 1, False), (start, 0, False), (start, 1, True), (start + 7, 1, True), (0, 1, False),
                                   (L - 2 * sps - 3, 1, False), (L, 0, False)):
        cfo = rng.uniform(-20, 20, B).astype(np.float32) if use_comp else None
        ph = rng.uniform(-3, 3, B).astype(np.float32) if use_comp else None
        if use_comp:
            cfo[0], ph[0] = 0.3, 0.005      # below both gates: no compensation
            cfo[1], ph[1] = 0.0, 0.02       # phase gate only
        n = dem.n_llr(L, ds)
        got = dem.demod_soft_batch(frames, ds, ref_mode, cfo, ph)
        for b in range(B):
            want = O.dpsk_demod_soft(mod, sps, frames[b], ds, ref_mode, 0.0 if cfo is None else float(cfo[b]), 0.0 if ph is None else float(ph[b]))
            assert len(want) == n
            assert same_bits(got[b, :n], want), (mod, sps, ds, ref_mode, use_comp, b)


@pytest.mark.parametrize("mod", [0, 1, 2])
def test_sc_dpsk_config4_frames_decode_like_the_oracle(ctx, mod):
    """SURVEY config 4: 125 baud, R1/4, 20-byte payload, Barker preamble + 648/324/216 data symbols; device memory."""
    import torch
    from projectultra_b200 import capi
    rng = np.random.default_rng(400 + mod)
    cfg = capi.dpsk_config(mod, 384)
    dem = capi.DpskDemodulator(ctx, cfg)
    dec = capi.LdpcDecoder(ctx, capi.R1_4)
    payload = rng.integers(0, 256, 20, dtype=np.uint8)
    tx = capi.dpsk_tx(cfg, capi.ldpc_encode(capi.R1_4, payload), 0)
    assert len(tx) == {0: 263808, 1: 139392, 2: 97920}[mod]
    tx = (tx * np.float32(0.5 / np.abs(tx).max())).astype(np.float32)
    snrs = {0: (-12.0, -9.0, -5.0), 1: (-8.0, -4.0, 0.0), 2: (-2.0, 2.0, 6.0)}[mod]
    frames = np.stack([noisy(tx, s, rng) for s in snrs for _ in range(4)])
    start = 39 * 384
    llr = dem.demod_soft_batch(torch.from_numpy(frames).cuda(), start, 1, llr_stride=648)
    info, ok, it = dec.decode_batch(llr)
    torch.cuda.synchronize()
    host = dem.demod_soft_batch(frames, start, 1, llr_stride=648)
    assert same_bits(host, llr.cpu().numpy())
    ref_llr = np.stack([O.dpsk_demod_soft(mod, 384, f, start, 1)[:648] for f in frames])
    assert same_bits(host, ref_llr)
    ci, cok, cit = O.ldpc_decode_batch(R.R1_4, ref_llr)
    assert (ok.cpu().numpy() == cok).all() and (it.cpu().numpy() == cit).all() and (info.cpu().numpy() == ci).all()
    good = ok.cpu().numpy().astype(bool)
    assert good[-4:].all() and (info.cpu().numpy()[good][:, :20] == payload).all()


@pytest.mark.parametrize("nc,bits", [(8, 2), (3, 2), (5, 2), (13, 2), (20, 2), (10, 1)])
def test_mc_dpsk_llrs_bit_identical(ctx, nc, bits):
    import torch
    from projectultra_b200 import capi
    rng = np.random.default_rng(7 * nc + bits)
    cfg = capi.mcdpsk_config(nc, bits)
    dem = capi.McDpskDemodulator(ctx, cfg)
    tx = capi.mcdpsk_tx(cfg, rng.integers(0, 256, 81, dtype=np.uint8))
    frames = np.stack([noisy(tx, snr, rng) for snr in (20.0, 6.0, 0.0, -8.0)] + [np.zeros_like(tx), tx * np.float32(2e-4)])
    B, L = frames.shape
    for Lcut in (L, L - 100, 9 * 512 + 700, 9 * 512):
        x = np.ascontiguousarray(frames[:, :Lcut])
        n = dem.n_llr(Lcut)
        got, cfo = dem.demod_soft_batch(x)
        dgot, dcfo = dem.demod_soft_batch(torch.from_numpy(x).cuda())
        torch.cuda.synchronize()
        assert same_bits(got, dgot.cpu().numpy()) and same_bits(cfo, dcfo.cpu().numpy())
        for b in range(B):
            want, wcfo = O.mcdpsk_demod_soft(nc, x[b], bits=bits)
            assert len(want) == n
            assert same_bits(got[b, :n], want), (nc, bits, Lcut, b)
            assert np.float32(cfo[b]).view(np.uint32) == np.float32(wcfo).view(np.uint32), (nc, b, cfo[b], wcfo)
    # one codeword through the decoder
    dec = capi.LdpcDecoder(ctx, capi.R1_2)
    payload = rng.integers(0, 256, 40, dtype=np.uint8)
    tx = capi.mcdpsk_tx(cfg, capi.ldpc_encode(capi.R1_2, payload))
    llr, _ = dem.demod_soft_batch(noisy(tx, 15.0, rng)[None, :], llr_stride=648)
    info, ok, it = dec.decode_batch(llr)
    assert ok[0] == 1 and (info[0, :40] == payload).all()


def test_golden_frames(ctx, golden):
    from projectultra_b200 import capi
    g = golden["psk"]
    for mod in (0, 1, 2):
        dem = capi.DpskDemodulator(ctx, capi.dpsk_config(mod, 192))
        rx = g[f"sc{mod}_rx"]
        assert same_bits(dem.demod_soft_batch(rx, 9 * 192, 1)[0], g[f"sc{mod}_llr_ref1"])
        assert same_bits(dem.demod_soft_batch(rx, 9 * 192, 0)[0], g[f"sc{mod}_llr_ref0"])
        comp = dem.demod_soft_batch(rx, 9 * 192, 1, np.array([7.25], np.float32), np.array([-0.6], np.float32))[0]
        assert same_bits(comp, g[f"sc{mod}_llr_comp"])
    for nc, bits in ((8, 2), (3, 2), (5, 1)):
        dem = capi.McDpskDemodulator(ctx, capi.mcdpsk_config(nc, bits))
        llr, cfo = dem.demod_soft_batch(g[f"mc{nc}_rx"])
        assert same_bits(llr[0], g[f"mc{nc}_llr"]) and cfo[0] == g[f"mc{nc}_cfo"][0]


@pytest.mark.parametrize("nc", [5, 8, 13, 20])
def test_mcdpsk_got_chirp_with_hilbert_cfo_correction(nc):
    """SURVEY §8a row a16, second half: MC-DPSK behind an externally detected chirp (MCDPSKWaveform::process = setChirpDetected
    -> process -> getSoftBits, i.e. processGotChirp): frames whose chirp CFO exceeds 0.1 Hz are first frequency-shifted through
    the 127-tap Hilbert FIR (real path delayed by 63 samples, Q16), then processTraining / the 5 Hz false-positive rule /
    setReference / demodulateSoft.  Soft-bit count and CFO report identical to the oracle, LLR words bit-identical (the
    corrected samples are, so everything behind them is); the compiled reference is checked on a subset."""
    import torch
    from projectultra_b200 import capi
    ctx = capi.Context(0)
    cfg = capi.mcdpsk_config(nc, 2)
    dem = capi.McDpskDemodulator(ctx, cfg)
    rng = np.random.default_rng(300 + nc)
    cases = [(20.0, 0.0), (12.0, 0.05), (12.0, 0.1), (10.0, 0.15), (8.0, -0.4), (6.0, 2.5), (15.0, -7.0), (25.0, 33.0), (3.0, 0.2), (30.0, -0.11)]
    frames = []
    for snr, _ in cases:
        tx = capi.mcdpsk_tx(cfg, capi.ldpc_encode(capi.R1_4, rng.integers(0, 256, 20, dtype=np.uint8)))
        p = float(np.mean(tx.astype(np.float64) ** 2))
        frames.append((tx + rng.normal(0.0, np.sqrt(p / 10 ** (snr / 10)), len(tx))).astype(np.float32))
    x = np.stack(frames)
    cfo = np.array([c for _, c in cases], np.float32)
    llr, n, after = dem.got_chirp_batch(x, cfo, llr_stride=700)
    accepted = 0
    for b in range(len(x)):
        ol, oafter = O.mcdpsk_got_chirp(nc, x[b], float(cfo[b]))
        want = ol[:700]
        assert int(n[b]) == len(want), (b, cases[b], n[b], len(ol))
        assert np.float32(after[b]).view(np.uint32) == np.float32(oafter).view(np.uint32), (b, after[b], oafter)
        if len(want):
            accepted += 1
            assert (llr[b, :len(want)].view(np.uint32) == want.view(np.uint32)).all(), (b, cases[b])
        if R.available() and b in (0, 3, 5):
            rl, rr, rc = R.mcdpsk_got_chirp(nc, x[b], float(cfo[b]))
            assert len(rl[:700]) == len(want) and (rl[:700].view(np.uint32) == want.view(np.uint32)).all()
    assert accepted >= 5
    d = dem.got_chirp_batch(torch.from_numpy(x).cuda(), cfo, llr_stride=700)
    torch.cuda.synchronize()
    assert (d[1].cpu().numpy() == n).all() and (d[0].cpu().numpy().view(np.uint32) == llr.view(np.uint32)).all()
    del ctx


@pytest.mark.parametrize("nc", [5, 8, 20])
def test_mcdpsk_chirp_receive_matches_oracle(nc):
    """SURVEY §8f next-2 on MC-DPSK: the whole IWaveform receive sequence of tools/test_iwaveform.cpp:127-160 behind the dual chirp
    (MCDPSKWaveform detectSync -> setFrequencyOffset -> process -> getSoftBits) for a ragged batch: detection flags, chirp positions,
    training start, CFO and correlation words, soft-bit count, CFO report and LLR words identical to the oracle (and to the compiled
    reference on a subset); covers CFO-shifted frames (Hilbert-FIR correction of the located span), a missed chirp, a frame cut
    inside its data and one cut right behind the preamble."""
    import torch
    from projectultra_b200 import capi
    from mcframes import mcdpsk_chirp_frame
    ctx = capi.Context(0)
    cfg = capi.mcdpsk_config(nc, 2)
    dem = capi.McDpskDemodulator(ctx, cfg)
    rng = np.random.default_rng(900 + nc)
    nsym = 9 + -(-648 // (2 * nc))
    total = 1500 + 57600 + nsym * 512 + 300
    cases = [(14.0, 500, 0.0, None), (6.0, 0, 7.3, None), (10.0, 1234, -22.0, None), (-14.0, 100, 0.0, None), (12.0, 300, 3.0, 60000 + 4096),
             (12.0, 0, 0.0, 57600 + 9 * 512), (9.0, 1500, -0.3, None), (20.0, 77, 0.08, None)]
    frames = []
    for snr, lead, cfo, cut in cases:
        f = mcdpsk_chirp_frame(cfg, rng, snr, lead, cfo, total)
        if cut is not None:       # a receiver buffer that ends early: the rest of the row is silence
            f = f.copy()
            f[cut:] = 0.0
        frames.append(f)
    x = np.stack(frames)
    llr, n, info, val, after = dem.chirp_receive_batch(x, llr_stride=700)
    detected = accepted = 0
    for b in range(len(x)):
        ol, oi, of, oa = O.mcdpsk_chirp_receive(nc, x[b])
        assert (info[b] == oi).all(), (b, cases[b], info[b], oi)
        assert (val[b, :3].view(np.uint32) == of.view(np.uint32)).all(), (b, val[b], of)
        want = ol[:700]
        assert int(n[b]) == len(want), (b, cases[b], n[b], len(ol))
        assert np.float32(after[b]).view(np.uint32) == np.float32(oa).view(np.uint32), (b, after[b], oa)
        detected += int(oi[0])
        if len(want):
            accepted += 1
            assert (llr[b, :len(want)].view(np.uint32) == want.view(np.uint32)).all(), (b, cases[b])
        if R.available() and b in (1, 3, 4):
            rl, ri, rf, ra = R.mcdpsk_chirp_receive(nc, x[b])
            assert (ri == oi).all() and len(rl[:700]) == len(want) and (rl[:700].view(np.uint32) == want.view(np.uint32)).all()
    assert detected >= 6 and accepted >= 3
    # buffers that end right behind the chirp pair: training start at or beyond the end (test_iwaveform.cpp:143) or no data behind the
    # preamble (processGotChirp keeps waiting) -> no soft bits, CFO report = the chirp's
    xs = np.ascontiguousarray(x[:, :58000])
    ls, ns, infos, vals, afters = dem.chirp_receive_batch(xs, llr_stride=700)
    for b in range(len(xs)):
        ol, oi, of, oa = O.mcdpsk_chirp_receive(nc, xs[b])
        assert (infos[b] == oi).all() and int(ns[b]) == len(ol) == 0, (b, infos[b], oi, ns[b], len(ol))
        assert np.float32(afters[b]).view(np.uint32) == np.float32(oa).view(np.uint32), (b, afters[b], oa)
    d = dem.chirp_receive_batch(torch.from_numpy(x).cuda(), llr_stride=700)
    torch.cuda.synchronize()
    assert (d[1].cpu().numpy() == n).all() and (d[2].cpu().numpy() == info).all()
    for b in range(len(x)):
        k = int(n[b])
        assert (d[0][b, :k].cpu().numpy().view(np.uint32) == llr[b, :k].view(np.uint32)).all()
    del ctx
