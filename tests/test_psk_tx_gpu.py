"""GPU transmitters of the PSK waveforms (csrc/psk_tx_gpu.cu: pu_dpsk_tx_batch, pu_mcdpsk_tx_batch) against the host transmitters
(pu_dpsk_tx / pu_mcdpsk_tx, pinned bit-identical to the unmodified reference's modulators by tests/test_psk_tx.py) and, when the
compiled reference is present, against the reference itself: every sample bit-identical, for every modulation / carrier count and
code rate, host and device memory, with and without the tools' peak normalisation; and the link simulator with a fresh payload per
trial (as tools/test_dpsk_snr.cpp draws one per trial) against the oracle on the same frames."""
import numpy as np
import pytest

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


@pytest.mark.parametrize("mod", [0, 1, 2])
def test_dpsk_tx_batch_is_bit_identical_to_the_host_transmitter(ctx, mod):
    import torch
    from projectultra_b200 import capi
    rng = np.random.default_rng(40 + mod)
    for sps, rate, nbytes in ((384, capi.R1_4, 20), (192, capi.R1_2, 40), (384, capi.R3_4, 60)):
        cfg = capi.dpsk_config(mod, sps)
        dem = capi.DpskDemodulator(ctx, cfg)
        enc = capi.LdpcDecoder(ctx, rate)
        payload = rng.integers(0, 256, (5, nbytes), dtype=np.uint8)
        want = np.stack([capi.dpsk_tx(cfg, capi.ldpc_encode(rate, p), 0) for p in payload])
        got = dem.tx_batch(enc, payload)
        assert got.shape == want.shape == (5, 39 * sps + sps * (648 // (mod + 1)))
        assert same_bits(got, want), (mod, sps, "host memory")
        dev = dem.tx_batch(enc, torch.from_numpy(payload).cuda())
        torch.cuda.synchronize()
        assert same_bits(dev.cpu().numpy(), want), (mod, sps, "device memory")
        peak = dem.tx_batch(enc, payload, peak=0.5)
        ref = np.stack([(w * (np.float32(0.5) / np.abs(w).max())).astype(np.float32) for w in want])
        assert same_bits(peak, ref), (mod, sps, "peak normalisation")
        if R.available():
            assert same_bits(got[0], R.dpsk_tx(mod, sps, R.ldpc_encode(rate, payload[0]), 0)), (mod, sps, "compiled reference")


@pytest.mark.parametrize("nc,bits", [(8, 2), (3, 2), (5, 1), (13, 2), (20, 2)])
def test_mcdpsk_tx_batch_is_bit_identical_to_the_host_transmitter(ctx, nc, bits):
    import torch
    from projectultra_b200 import capi
    rng = np.random.default_rng(70 + nc)
    cfg = capi.mcdpsk_config(nc, bits)
    dem = capi.McDpskDemodulator(ctx, cfg)
    for rate, nbytes in ((capi.R1_4, 20), (capi.R1_2, 40), (capi.R5_6, 67)):
        enc = capi.LdpcDecoder(ctx, rate)
        payload = rng.integers(0, 256, (4, nbytes), dtype=np.uint8)
        want = np.stack([capi.mcdpsk_tx(cfg, capi.ldpc_encode(rate, p)) for p in payload])
        got = dem.tx_batch(enc, payload)
        assert same_bits(got, want), (nc, bits, rate, "host memory")
        dev = dem.tx_batch(enc, torch.from_numpy(payload).cuda(), peak=0.5)
        torch.cuda.synchronize()
        ref = np.stack([(w * (np.float32(0.5) / np.abs(w).max())).astype(np.float32) for w in want])
        assert same_bits(dev.cpu().numpy(), ref), (nc, bits, rate, "device memory + peak")
        if R.available():
            assert same_bits(got[0], R.mcdpsk_tx(nc, R.ldpc_encode(rate, payload[0]), bits=bits)), (nc, bits, "compiled reference")


@pytest.mark.parametrize("kind", ["dpsk", "mcdpsk"])
def test_linksim_with_a_fresh_payload_per_trial(ctx, kind):
    """LinkSim(fresh_payload=True) on the PSK waveforms: payload -> GPU transmitter -> channel -> demod -> LDPC -> counters, with the
    frames regenerated on the host (host transmitter + channel twin) and decoded by the oracle: identical verdicts and counters."""
    import torch
    import channelapi as CH
    import oracleapi as O
    from projectultra_b200 import capi, linksim
    cfg = capi.dpsk_config(1, 384) if kind == "dpsk" else capi.mcdpsk_config(8, 2)
    rate = capi.R1_4 if kind == "dpsk" else capi.R1_2
    nbytes = 20 if kind == "dpsk" else 40
    sim = linksim.LinkSim(ctx, cfg, "poor" if kind == "dpsk" else "good", payload_bytes=nbytes, pool=2, code_rate=rate, peak=0.5, fresh_payload=True)
    snrs = [-2.0, 4.0, 10.0]
    trials = 6
    si = np.repeat(np.arange(len(snrs)), trials)
    tr = np.tile(np.arange(trials), len(snrs))
    batch = sim.make_batch(snrs, si, tr)
    counters = torch.zeros((len(snrs), 6), dtype=torch.int64, device="cuda")
    rx, info, ok, iters = sim.run_batch(batch, counters, keep=True, snr_points=snrs)
    torch.cuda.synchronize()
    payload = sim.last_payload.cpu().numpy()
    tx = sim.last_tx.cpu().numpy()
    std = sim.last_std.cpu().numpy()
    rx_h = rx.cpu().numpy()
    frame_err = np.zeros(len(snrs), np.int64)
    for b in range(len(si)):
        coded = capi.ldpc_encode(rate, payload[b, :nbytes])
        w = capi.dpsk_tx(cfg, coded, 0) if kind == "dpsk" else capi.mcdpsk_tx(cfg, coded)
        w = (w * (np.float32(0.5) / np.abs(w).max())).astype(np.float32)
        assert same_bits(tx[b], w), b
        assert same_bits(rx_h[b], CH.channel_apply(sim.ch, w, std[b], batch["host"]["seed"][b])), b
    if kind == "dpsk":
        llr = O.dpsk_demod_soft_batch(cfg, rx_h, sim.data_start, 1)[:, :648] if hasattr(O, "dpsk_demod_soft_batch") else None
    else:
        llr = None
    if llr is not None:
        cinfo, cok, cit = O.ldpc_decode_batch(rate, np.ascontiguousarray(llr))
        assert (ok.cpu().numpy() == cok).all() and (iters.cpu().numpy() == cit).all()
    okh, infoh = ok.cpu().numpy(), info.cpu().numpy()
    for b in range(len(si)):
        good = okh[b] == 1 and (infoh[b, :nbytes] == payload[b, :nbytes]).all()
        frame_err[si[b]] += 0 if good else 1
    c = counters.cpu().numpy()
    assert (c[:, 0] == trials).all() and (c[:, 1] == frame_err).all()
    assert c[-1, 1] == 0          # 10 dB: every frame decodes
