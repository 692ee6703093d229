"""Schmidl-Cox acquisition and the process() path on the GPU against the unmodified reference
(oracle/_ref: OFDMDemodulator::process fed in 960-sample chunks exactly as tools/test_mode_snr.cpp:65-70 does, then
getSoftBits()).  Integer results (synchronised or not, sync offset, samples consumed) must be identical, the coarse CFO
and every LLR bit-identical up to the tolerance of the other OFDM tests (1e-4 relative)."""
import numpy as np
import pytest

import refapi as R
import oracleapi as O

pytestmark = pytest.mark.gpu


def awgn(tx, snr_db, rng):
    p = float(np.mean(tx.astype(np.float64) ** 2))       # tools/test_mode_snr.cpp:58-63: mean frame power
    return (tx + rng.normal(0.0, np.sqrt(p / 10 ** (snr_db / 10)), len(tx))).astype(np.float32)


def sc_frame(cfg, rate, nbytes, snr, seed, lead=0, tail=0):
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, nbytes, dtype=np.uint8)
    tx = O.ofdm_tx(cfg, O.ldpc_encode(rate, data), 1)     # layout 1: generatePreamble() + modulate() (Schmidl-Cox frame)
    tx = np.concatenate([np.zeros(lead, np.float32), tx, np.zeros(tail, np.float32)])
    return awgn(tx, snr, rng) if snr is not None else tx.astype(np.float32)


@pytest.mark.parametrize("preset,mod", [("m1", R.DQPSK), ("m1", R.D8PSK), ("m1", R.DBPSK), ("m3", R.DQPSK)])
def test_process_path_matches_reference(preset, mod):
    if not R.available():
        pytest.skip("needs oracle/_ref (the compiled reference)")
    from projectultra_b200 import capi
    rate = R.R1_2 if preset == "m1" else R.R3_4
    cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
    nbytes = 40 if preset == "m1" else 60
    ctx = capi.Context(0)
    dem = capi.OfdmDemodulator(ctx, capi.ModemConfig.from_buffer_copy(bytes(cfg)))
    cases = [(30.0, 0), (25.0, 0), (22.0, 0), (18.0, 0), (14.0, 0), (8.0, 0), (None, 0), (28.0, 0), (26.0, 0), (20.0, 0), (16.0, 0), (12.0, 0)]
    frames = [sc_frame(cfg, rate, nbytes, snr, 500 + 17 * i + mod) for i, (snr, _) in enumerate(cases)]
    frames.append(np.zeros_like(frames[0]))                                            # silence: never synchronises
    frames.append(np.random.default_rng(3).normal(0, 0.1, len(frames[0])).astype(np.float32))   # noise only
    x = np.stack(frames)
    for chunk in (960, 4096, x.shape[1]):
        llr, n, info, cfo, snr = dem.process_batch(x, chunk=chunk)
        info2, cfo2 = dem.acquire_batch(x, chunk=chunk)
        assert (info2 == info).all() and (cfo2.view(np.uint32) == cfo.view(np.uint32)).all()
        n_sync = 0
        for b in range(len(x)):
            rl, rs, roff, rcfo = R.ofdm_process_info(cfg, x[b], chunk)
            assert bool(info[b, 0]) == rs, (b, chunk, info[b], rs, roff)
            if not rs:
                assert n[b] == 0 and not llr[b].any()
                continue
            n_sync += 1
            assert int(info[b, 1]) == roff, (b, chunk, info[b], roff)
            assert np.float32(cfo[b]).view(np.uint32) == np.float32(rcfo).view(np.uint32), (b, chunk, cfo[b], rcfo)
            assert int(n[b]) == len(rl), (b, chunk, n[b], len(rl), info[b])
            got, want = llr[b, :len(rl)], rl
            bad = np.flatnonzero(~np.isclose(got, want, rtol=1e-4, atol=1e-6))
            assert len(bad) == 0, (b, chunk, bad[:8], got[bad[:8]], want[bad[:8]])
        assert n_sync >= 3, "the case list must contain frames the reference synchronises on"
    del ctx


def test_golden_and_oracle_acquisition(golden):
    """The same path against the committed reference vectors (tests/golden/acquire_golden.npz) and against the plain-C oracle
    (oracle/pu_oracle_ofdm.c: orc_ofdm_process) on a batch that mixes synchronising and non-synchronising frames and carries
    leading silence, so that the energy gate, the noise-floor tracker and later process() calls are exercised."""
    from golden.make_golden import ACQ_CASES
    from projectultra_b200 import capi
    ctx = capi.Context(0)
    g = golden["acquire"]
    for i in range(len(ACQ_CASES)):
        cfg = R.ModemConfig.from_buffer_copy(bytes(g[f"a{i}_cfg"]))
        dem = capi.OfdmDemodulator(ctx, capi.ModemConfig.from_buffer_copy(bytes(cfg)))
        rx = g[f"a{i}_rx"]
        for ci, chunk in enumerate((960, len(rx))):
            llr, n, info, cfo, _ = dem.process_batch(rx[None, :], chunk=chunk)
            want_sync, want_off = (int(v) for v in g[f"a{i}_c{ci}_info"])
            assert int(info[0, 0]) == want_sync, (i, chunk)
            if want_sync:
                want = g[f"a{i}_c{ci}_llr"]
                assert int(info[0, 1]) == want_off and cfo.view(np.uint32)[0] == g[f"a{i}_c{ci}_cfo"].view(np.uint32)[0]
                assert int(n[0]) == len(want) and np.allclose(llr[0, :len(want)], want, rtol=1e-4, atol=1e-6), (i, chunk)
    cfg = R.config_m1(R.DQPSK, R.R1_2)
    dem = capi.OfdmDemodulator(ctx, capi.ModemConfig.from_buffer_copy(bytes(cfg)))
    L = 10124 + 3000
    frames = []
    for i, (snr, lead) in enumerate(((30.0, 0), (26.0, 1000), (22.0, 2999), (19.0, 1777), (16.0, 500), (13.0, 0), (6.0, 2000), (24.0, 8))):
        f = sc_frame(cfg, R.R1_2, 40, snr, 4000 + i, lead=lead, tail=3000 - lead)
        frames.append(f + np.random.default_rng(i).normal(0, 1e-3, len(f)).astype(np.float32))
    x = np.stack(frames)
    assert x.shape[1] == L
    for chunk in (960, 1500):
        llr, n, info, cfo, _ = dem.process_batch(x, chunk=chunk)
        n_sync = 0
        for b in range(len(x)):
            ol, osync, ooff, ocfo, ods, ocalls = O.ofdm_process(cfg, x[b], chunk)
            assert bool(info[b, 0]) == osync, (b, chunk, info[b])
            if osync:
                n_sync += 1
                assert (int(info[b, 1]), int(info[b, 2]), int(info[b, 3])) == (ooff, ods, ocalls), (b, chunk, info[b], ooff, ods, ocalls)
                assert np.float32(cfo[b]).view(np.uint32) == np.float32(ocfo).view(np.uint32)
                assert int(n[b]) == len(ol) and np.allclose(llr[b, :len(ol)], ol, rtol=1e-4, atol=1e-6), (b, chunk)
        assert n_sync >= 3
    del ctx


def test_acquire_device_and_host_paths_agree():
    import torch
    from projectultra_b200 import capi
    cfg = R.config_m1(R.DQPSK, R.R1_2)
    ctx = capi.Context(0)
    dem = capi.OfdmDemodulator(ctx, capi.ModemConfig.from_buffer_copy(bytes(cfg)))
    x = np.stack([sc_frame(cfg, R.R1_2, 40, snr, 90 + i) for i, snr in enumerate((30.0, 24.0, 19.0, 9.0))])
    h_llr, h_n, h_info, h_cfo, _ = dem.process_batch(x)
    d_llr, d_n, d_info, d_cfo, _ = dem.process_batch(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    assert (d_info.cpu().numpy() == h_info).all() and (d_n.cpu().numpy() == h_n).all()
    assert (d_llr.cpu().numpy().view(np.uint32) == h_llr.view(np.uint32)).all()
    assert (d_cfo.cpu().numpy().view(np.uint32) == h_cfo.view(np.uint32)).all()
    assert h_info[0, 0] == 1 and h_n[0] == 648
    with pytest.raises(capi.PuError):
        dem.acquire_batch(np.zeros((1, 40001), np.float32))
    del ctx
