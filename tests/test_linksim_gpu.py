"""End-to-end parity of the batched link (channel -> demod -> LDPC -> counters) against the oracle on IDENTICAL
channel outputs: every frame's success flag, iteration count and decoded bytes, hence identical FER/BER counters;
frames regenerate bit-for-bit on the CPU from (waveform, sigma, seed); sharding over ranks does not change totals."""
import numpy as np
import pytest

import channelapi as CH
import oracleapi as O
import refapi as R

pytestmark = pytest.mark.gpu


def to_capi_cfg(cfg):
    from projectultra_b200 import capi
    return capi.ModemConfig.from_buffer_copy(bytes(cfg))


def cpu_counters(cfg, rate, rx, payloads, tx_index, snr_idx, n_snr, payload_bytes):
    llr, counts = O.ofdm_presynced_batch(cfg, rx, 648)
    info, ok, it = O.ldpc_decode_batch(rate, llr)
    c = np.zeros((n_snr, 6), np.int64)
    for b in range(len(rx)):
        want = payloads[tx_index[b]]
        berr = int(np.unpackbits(info[b, :payload_bytes] ^ want).sum())
        succ = bool(ok[b]) and berr == 0
        r = c[snr_idx[b]]
        r[0] += 1
        r[1] += 0 if succ else 1
        r[2] += berr
        r[3] += payload_bytes * 8
        r[4] += 0 if ok[b] else 1
        r[5] += int(it[b])
    return c, info, ok, it


@pytest.mark.parametrize("case", [("m1", R.DQPSK, R.R1_2, 40, "awgn", (-3.0, -1.0, 0.0, 1.0, 3.0), 48),
                                  ("m1", R.QAM16, R.R1_2, 40, "moderate", (8.0, 14.0, 20.0), 24),
                                  ("m3", R.QAM32, R.R3_4, 60, "good", (8.0, 12.0, 16.0, 24.0), 24),
                                  ("m1", R.D8PSK, R.R1_4, 20, "flutter", (0.0, 6.0, 12.0), 24)])
def test_frame_by_frame_parity(case):
    import torch
    from projectultra_b200 import capi, linksim
    preset, mod, rate, nbytes, chan, snrs, trials = case
    cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
    ctx = capi.Context(0)
    sim = linksim.LinkSim(ctx, to_capi_cfg(cfg), chan, payload_bytes=nbytes, pool=8)
    si = np.repeat(np.arange(len(snrs)), trials)
    tr = np.tile(np.arange(trials), len(snrs))
    batch = sim.make_batch(snrs, si, tr)
    counters = torch.zeros((len(snrs), 6), dtype=torch.int64, device="cuda")
    rx, info, ok, iters = sim.run_batch(batch, counters, keep=True)
    torch.cuda.synchronize()
    rx_h = rx.cpu().numpy()
    h = batch["host"]
    # (1) frames regenerate on the CPU from (waveform, sigma, seed)
    for b in (0, 7, len(si) - 1):
        twin = CH.channel_apply(sim.ch, sim.tx_host[h["tx_index"][b]], h["noise_std"][b], h["seed"][b])
        assert (twin.view(np.uint32) == rx_h[b].view(np.uint32)).all()
    # (2) oracle on the same channel outputs
    want, cinfo, cok, cit = cpu_counters(cfg, rate, rx_h, sim.payloads, h["tx_index"], si, len(snrs), nbytes)
    assert (ok.cpu().numpy() == cok).all()
    assert (iters.cpu().numpy() == cit).all()
    assert (info.cpu().numpy() == cinfo).all()
    assert (counters.cpu().numpy() == want).all()
    fer = want[:, 1] / want[:, 0]
    assert fer[0] >= fer[-1]                      # waterfall runs the right way
    del ctx


def test_sharding_invariance_and_waterfall():
    import torch
    from projectultra_b200 import capi, linksim
    cfg = R.config_m1(R.DQPSK, R.R1_2)
    ctx = capi.Context(0)
    sim = linksim.LinkSim(ctx, to_capi_cfg(cfg), "awgn", payload_bytes=40, pool=16)
    snrs = [-4.0, -2.0, 0.0, 2.0, 4.0]
    whole = sim.sweep(snrs, 300, batch_frames=700)
    parts = sum(sim.sweep(snrs, 300, rank=r, world=3, batch_frames=256) for r in range(3))
    torch.cuda.synchronize()
    assert (whole == parts).all()
    rows = linksim.summarize(whole, snrs)
    # SURVEY §8d [probe]: M1 DQPSK R1/2 success 0 % @-4 dB, ~68 % @0 dB, 100 % @+4 dB (wideband mean-power SNR)
    assert rows[0]["fer"] > 0.9 and rows[-1]["fer"] < 0.02 and 0.1 < rows[2]["fer"] < 0.7
    assert all(r["frames"] == 300 for r in rows)
    del ctx


def test_host_buffer_path_matches_device_path():
    import torch
    from projectultra_b200 import capi, linksim
    cfg = R.config_m1(R.DQPSK, R.R1_2)
    ctx = capi.Context(0)
    sim = linksim.LinkSim(ctx, to_capi_cfg(cfg), "awgn", payload_bytes=40, pool=4)
    batch = sim.make_batch([0.0, 5.0], np.repeat([0, 1], 40), np.tile(np.arange(40), 2))
    rx = linksim.channel_apply(ctx, sim.ch, sim.tx_pool, batch["tx_index"], batch["noise_std"], batch["seed"])
    a = linksim.receive_decode(sim.ofdm, sim.ldpc, rx)
    torch.cuda.synchronize()
    b = linksim.receive_decode(sim.ofdm, sim.ldpc, rx.cpu().numpy())
    for u, v in zip(a, b):
        assert (u.cpu().numpy() == v).all()
    # pinned host buffers are DMA'd in place, and for zero-CFO 512-FFT differential frames only the FFT windows of the symbols behind
    # the first training symbol cross PCIe (12 x 512 of 7 332 samples): poison everything else -- results must not change
    B, L = rx.shape
    pinned = torch.empty((B, L), dtype=torch.float32, pin_memory=True)
    pinned.copy_(rx)
    t0 = ctx.transfer_bytes
    c = linksim.receive_decode(sim.ofdm, sim.ldpc, pinned.numpy())
    t1 = ctx.transfer_bytes
    for u, v in zip(a, c):
        assert (u.cpu().numpy() == v).all()
    assert t1[0] - t0[0] == B * 12 * 512 * 4 and t1[1] > t0[1]
    poisoned = pinned.clone().view(B, 13, 564)
    poisoned[:, 0, :] = float("nan")
    poisoned[:, :, :48] = float("nan")
    poisoned[:, :, 560:] = float("nan")
    pp = torch.empty((B, L), dtype=torch.float32, pin_memory=True)
    pp.copy_(poisoned.view(B, L))
    d = linksim.receive_decode(sim.ofdm, sim.ldpc, pp.numpy())
    for u, v in zip(a, d):
        assert (u.cpu().numpy() == v).all()
    # a call with a CFO takes the general kernel and copies whole frames
    t0 = ctx.transfer_bytes
    linksim.receive_decode(sim.ofdm, sim.ldpc, pinned.numpy(), cfo_hz=np.full(B, 0.5, np.float32))
    assert ctx.transfer_bytes[0] - t0[0] >= B * L * 4
    del ctx


@pytest.mark.parametrize("case", [("dpsk", 1, R.R1_4, 20, "poor", (-6.0, 0.0, 8.0), 6),
                                  ("dpsk", 0, R.R1_4, 20, "flutter", (-11.0, -3.0, 6.0), 4),
                                  ("dpsk", 2, R.R1_2, 40, "awgn", (2.0, 8.0, 17.0), 6),
                                  ("mcdpsk", 8, R.R1_2, 40, "good", (-2.0, 4.0, 12.0), 12),
                                  ("mcdpsk", 20, R.R1_4, 20, "moderate", (-6.0, 0.0, 8.0), 12)])
def test_psk_waveforms_frame_by_frame_parity(case):
    """SURVEY configs 4 and 5: single- and multi-carrier DPSK through channel -> demod -> LDPC -> counters, every frame's
    LLR words, ok flag, iteration count and bytes identical to the oracle on identical channel outputs."""
    import torch
    from projectultra_b200 import capi, linksim
    kind, p0, rate, nbytes, chan, snrs, trials = case
    cfg = capi.dpsk_config(p0, 384) if kind == "dpsk" else capi.mcdpsk_config(p0, 2)
    ctx = capi.Context(0)
    sim = linksim.LinkSim(ctx, cfg, chan, payload_bytes=nbytes, pool=4, code_rate=rate, peak=0.5 if kind == "dpsk" else None)
    si = np.repeat(np.arange(len(snrs)), trials)
    tr = np.tile(np.arange(trials), len(snrs))
    batch = sim.make_batch(snrs, si, tr)
    counters = torch.zeros((len(snrs), 6), dtype=torch.int64, device="cuda")
    rx, info, ok, iters = sim.run_batch(batch, counters, keep=True)
    llr = sim.demod_llr(rx)
    torch.cuda.synchronize()
    rx_h = rx.cpu().numpy()
    h = batch["host"]
    twin = CH.channel_apply(sim.ch, sim.tx_host[h["tx_index"][1]], h["noise_std"][1], h["seed"][1])
    assert (twin.view(np.uint32) == rx_h[1].view(np.uint32)).all()
    if kind == "dpsk":
        ref_llr = np.stack([O.dpsk_demod_soft(p0, 384, f, sim.data_start, 1)[:648] for f in rx_h])
    else:
        ref_llr = np.stack([O.mcdpsk_demod_soft(p0, f, bits=2)[0][:648] for f in rx_h])
    assert (llr.cpu().numpy().view(np.uint32) == ref_llr.view(np.uint32)).all()
    cinfo, cok, cit = O.ldpc_decode_batch(rate, ref_llr)
    assert (ok.cpu().numpy() == cok).all() and (iters.cpu().numpy() == cit).all() and (info.cpu().numpy() == cinfo).all()
    c = counters.cpu().numpy()
    assert c[:, 0].tolist() == [trials] * len(snrs)
    assert c[0, 1] >= c[-1, 1] and c[-1, 1] < trials        # the waterfall runs the right way and the top point decodes
    del ctx


def test_count_errors_matches_numpy_for_many_bins():
    """pu_count_errors: the frame-error rule of tools/test_mode_snr.cpp:98-104 over random decoder outputs, with more
    bins than the kernel's shared-memory table holds (bins >= 64 go straight to global memory) and ragged batch sizes."""
    import torch
    from projectultra_b200 import capi, linksim
    ctx = capi.Context(0)
    rng = np.random.default_rng(7)
    for B, n_bins in ((1, 1), (1000, 13), (5003, 100), (70000, 64)):
        nbytes, kb, pool = 40, 41, 8
        payloads = rng.integers(0, 256, (pool, nbytes), dtype=np.uint8)
        tx = rng.integers(0, pool, B).astype(np.uint32)
        info = np.zeros((B, kb), np.uint8)
        info[:, :nbytes] = payloads[tx]
        bad = rng.random(B) < 0.3
        info[bad, rng.integers(0, nbytes)] ^= rng.integers(1, 256, int(bad.sum()), dtype=np.uint8)
        ok = (rng.random(B) < 0.8).astype(np.uint8)
        iters = rng.integers(0, 51, B).astype(np.int32)
        bins = rng.integers(0, n_bins, B).astype(np.uint32)
        want = np.zeros((n_bins, 6), np.int64)
        biterr = np.unpackbits(info[:, :nbytes] ^ payloads[tx], axis=1).sum(axis=1)
        np.add.at(want[:, 0], bins, 1)
        np.add.at(want[:, 1], bins, ((ok == 0) | (biterr > 0)).astype(np.int64))
        np.add.at(want[:, 2], bins, biterr)
        np.add.at(want[:, 3], bins, nbytes * 8)
        np.add.at(want[:, 4], bins, (ok == 0).astype(np.int64))
        np.add.at(want[:, 5], bins, iters)
        t = lambda a: torch.from_numpy(a).cuda()
        counters = torch.zeros((n_bins, 6), dtype=torch.int64, device="cuda")
        for _ in range(2):
            linksim.count_errors(ctx, t(info), t(ok), t(iters), t(payloads), t(tx), t(bins), nbytes, counters)
        torch.cuda.synchronize()
        assert (counters.cpu().numpy() == 2 * want).all(), (B, n_bins)
    del ctx


def test_linksim_config1_with_acquisition():
    """BASELINE config 1 as tools/test_mode_snr.cpp runs it: frame = generatePreamble() + modulate(), peak-normalised to 0.5,
    AWGN on mean frame power, process() in 960-sample chunks (Schmidl-Cox acquisition), getSoftBits, decodeSoft, frame-error
    rule -- every frame's sync decision, soft-bit count, ok flag, iteration count and bytes against the oracle on the
    identical channel outputs."""
    import torch
    from projectultra_b200 import capi, linksim
    cfg = R.config_m1(R.DQPSK, R.R1_2)
    ctx = capi.Context(0)
    sim = linksim.LinkSim(ctx, capi.ModemConfig.from_buffer_copy(bytes(cfg)), "awgn", payload_bytes=40, pool=4, peak=0.5, layout="sc")
    assert sim.L == 10124
    snrs = [12.0, 16.0, 18.0, 21.0, 25.0]
    trials = 6
    si = np.repeat(np.arange(len(snrs)), trials)
    tr = np.tile(np.arange(trials), len(snrs))
    batch = sim.make_batch(snrs, si, tr)
    counters = torch.zeros((len(snrs), 6), dtype=torch.int64, device="cuda")
    rx, info, ok, iters = sim.run_batch(batch, counters, keep=True)
    torch.cuda.synchronize()
    rx_h = rx.cpu().numpy()
    n_llr = sim.last_n_llr.cpu().numpy()
    sync = sim.last_sync.cpu().numpy()
    ok_h, info_h, it_h = ok.cpu().numpy(), info.cpu().numpy(), iters.cpu().numpy()
    for b in range(len(rx_h)):
        ol, osync, ooff, ocfo, ods, ocalls = O.ofdm_process(cfg, rx_h[b], 960)
        assert bool(sync[b, 0]) == osync and int(n_llr[b]) == len(ol), (b, sync[b], n_llr[b], len(ol))
        if len(ol) >= 648:
            ci, cok, cit = O.ldpc_decode_batch(R.R1_2, ol[None, :648].copy())
            assert ok_h[b] == cok[0] and it_h[b] == cit[0] and (info_h[b] == ci[0]).all(), b
        else:
            assert ok_h[b] == 0
    c = counters.cpu().numpy()
    assert c[:, 0].tolist() == [trials] * len(snrs)
    assert c[0, 1] == trials and c[-1, 1] == 0      # the reference's acquisition threshold sits between 12 and 25 dB
    del ctx


@pytest.mark.parametrize("layout", ["presynced", "sc"])
def test_linksim_fresh_payload_gpu_transmitter(layout):
    """Every frame carries its own payload, LDPC-encoded and modulated on the GPU: the transmitted
    waveforms are the host/oracle transmitter's bit for bit, the noise level follows the tools' convention, the channel
    output regenerates on the CPU, and every frame's decode result matches the oracle on the identical channel output."""
    import torch
    from projectultra_b200 import capi, linksim
    cfg = R.config_m1(R.DQPSK, R.R1_2)
    ctx = capi.Context(0)
    sim = linksim.LinkSim(ctx, capi.ModemConfig.from_buffer_copy(bytes(cfg)), "awgn", payload_bytes=40, pool=2,
                          peak=0.5 if layout == "sc" else None, layout=layout, fresh_payload=True)
    snrs = [0.0, 3.0, 24.0] if layout == "presynced" else [14.0, 20.0, 26.0]
    trials = 5
    si = np.repeat(np.arange(len(snrs)), trials)
    tr = np.tile(np.arange(trials), len(snrs))
    batch = sim.make_batch(snrs, si, tr)
    counters = torch.zeros((len(snrs), 6), dtype=torch.int64, device="cuda")
    rx, info, ok, iters = sim.run_batch(batch, counters, keep=True, snr_points=snrs)
    torch.cuda.synchronize()
    payload = sim.last_payload.cpu().numpy()
    tx, std, rx_h = sim.last_tx.cpu().numpy(), sim.last_std.cpu().numpy(), rx.cpu().numpy()
    assert len(np.unique(payload[:, :40], axis=0)) == len(payload), "payloads must differ from frame to frame"
    h = batch["host"]
    for b in range(len(tx)):
        w = O.ofdm_tx(cfg, O.ldpc_encode(R.R1_2, payload[b, :40]), 1 if layout == "sc" else 0)
        if layout == "sc":
            w = (w * (np.float32(0.5) / np.abs(w).max())).astype(np.float32)
        assert (w.view(np.uint32) == tx[b].view(np.uint32)).all(), b
        assert abs(std[b] - CH.noise_std(w, snrs[si[b]], 1)) <= 2e-7 * std[b], (b, std[b])
        twin = CH.channel_apply(sim.ch, w, std[b], h["seed"][b])
        assert (twin.view(np.uint32) == rx_h[b].view(np.uint32)).all(), b
    if layout == "presynced":
        llr, _ = O.ofdm_presynced_batch(cfg, rx_h, 648)
        cinfo, cok, cit = O.ldpc_decode_batch(R.R1_2, llr)
    else:
        llr = np.zeros((len(rx_h), 648), np.float32)
        got = np.zeros(len(rx_h), bool)
        for b in range(len(rx_h)):
            ol = O.ofdm_process(cfg, rx_h[b], 960)[0]
            got[b] = len(ol) >= 648
            llr[b, :len(ol)] = ol
        cinfo, cok, cit = O.ldpc_decode_batch(R.R1_2, llr)
        cok = cok * got
    assert (ok.cpu().numpy() == cok).all() and (iters.cpu().numpy() == cit).all() and (info.cpu().numpy() == cinfo).all()
    c = counters.cpu().numpy()
    want_err = [(1 - ((cok == 1) & (cinfo[:, :40] == payload[:, :40]).all(axis=1))[si == k]).sum() for k in range(len(snrs))]
    assert c[:, 0].tolist() == [trials] * len(snrs) and c[:, 1].tolist() == want_err
    assert c[-1, 1] == 0
    del ctx


def test_linksim_config4_with_barker_acquisition():
    """BASELINE config 4 as tools/test_dpsk_snr.cpp runs it: Barker preamble + data, peak-normalised, findPreamble on the whole
    frame, demodulateSoft from the returned data start, decodeSoft, frame-error rule -- over the Watterson 'poor' channel, every
    frame's data start, soft bits, ok flag, iteration count and bytes against the oracle on the identical channel outputs."""
    import torch
    from projectultra_b200 import capi, linksim
    ctx = capi.Context(0)
    sim = linksim.LinkSim(ctx, capi.dpsk_config(1, 384), "poor", payload_bytes=20, pool=3, code_rate=capi.R1_4, peak=0.5, acquire=True)
    snrs = [-14.0, -4.0, 6.0, 16.0]
    trials = 4
    si = np.repeat(np.arange(len(snrs)), trials)
    tr = np.tile(np.arange(trials), len(snrs))
    batch = sim.make_batch(snrs, si, tr)
    counters = torch.zeros((len(snrs), 6), dtype=torch.int64, device="cuda")
    rx, info, ok, iters = sim.run_batch(batch, counters, keep=True)
    torch.cuda.synchronize()
    rx_h = rx.cpu().numpy()
    n_llr, ds = sim.last_n_llr.cpu().numpy(), sim.last_sync.cpu().numpy()
    ok_h, info_h, it_h = ok.cpu().numpy(), info.cpu().numpy(), iters.cpu().numpy()
    for b in range(len(rx_h)):
        ol, ods, ocfo, oph = O.dpsk_receive(1, 384, rx_h[b])
        assert int(ds[b]) == ods and int(n_llr[b]) == min(len(ol), 648), (b, ds[b], ods)
        if len(ol) >= 648:
            ci, cok, cit = O.ldpc_decode_batch(R.R1_4, ol[None, :648].copy())
            assert ok_h[b] == cok[0] and it_h[b] == cit[0] and (info_h[b] == ci[0]).all(), b
        else:
            assert ok_h[b] == 0
    c = counters.cpu().numpy()
    assert c[:, 0].tolist() == [trials] * len(snrs) and c[0, 1] >= c[-1, 1]
    del ctx


@pytest.mark.parametrize("chan", ["awgn", "good"])
def test_linksim_ofdm_chirp_with_dual_chirp_sync(chan):
    """OFDM_CHIRP: dual chirp + training + DQPSK data over the Watterson 'good' channel, received as
    tools/test_iwaveform.cpp:127-160 does (detectSync -> setFrequencyOffset -> process -> getSoftBits -> decodeSoft), every frame's
    detection result, soft-bit count, ok flag, iteration count and bytes against the oracle on the identical channel outputs."""
    import torch
    from projectultra_b200 import capi, linksim
    cfg = R.config_m1(R.DQPSK, R.R1_2)
    ctx = capi.Context(0)
    sim = linksim.LinkSim(ctx, capi.ModemConfig.from_buffer_copy(bytes(cfg)), chan, payload_bytes=40, pool=3, layout="chirp")
    assert sim.L == 57600 + 7332
    snrs = [-6.0, 2.0, 10.0, 22.0]
    trials = 3
    si = np.repeat(np.arange(len(snrs)), trials)
    tr = np.tile(np.arange(trials), len(snrs))
    batch = sim.make_batch(snrs, si, tr)
    counters = torch.zeros((len(snrs), 6), dtype=torch.int64, device="cuda")
    rx, info, ok, iters = sim.run_batch(batch, counters, keep=True)
    torch.cuda.synchronize()
    rx_h = rx.cpu().numpy()
    n_llr, sync = sim.last_n_llr.cpu().numpy(), sim.last_sync.cpu().numpy()
    ok_h, info_h, it_h = ok.cpu().numpy(), info.cpu().numpy(), iters.cpu().numpy()
    for b in range(len(rx_h)):
        ol, oi, ocfo = O.ofdm_chirp_receive(cfg, rx_h[b])
        assert (sync[b] == oi).all() and int(n_llr[b]) == min(len(ol), 648), (b, sync[b], oi, n_llr[b], len(ol))
        if len(ol) >= 648:
            ci, cok, cit = O.ldpc_decode_batch(R.R1_2, ol[None, :648].copy())
            # LLRs agree to 1e-4 (rotator path); decisions must agree on frames decoded with margin
            assert ok_h[b] == cok[0] and (cok[0] == 0 or (it_h[b] == cit[0] and (info_h[b] == ci[0]).all())), b
        else:
            assert ok_h[b] == 0
    c = counters.cpu().numpy()
    assert c[:, 0].tolist() == [trials] * len(snrs)
    # over two equal-gain paths the reference's chirp timing can lock on the later path (late FFT window => inter-symbol
    # interference), so only the AWGN run is required to be error free at 22 dB; the fading run checks parity frame by frame
    assert chan != "awgn" or c[-1, 1] == 0
    del ctx


@pytest.mark.parametrize("case", [(8, "awgn", (-8.0, 0.0, 6.0, 14.0)), (13, "good", (-4.0, 4.0, 12.0, 20.0))])
def test_linksim_mcdpsk_behind_dual_chirp(case):
    """MC-DPSK as transmitted: dual chirp + training + reference + DQPSK data, R1/4, received as
    tools/test_iwaveform.cpp:127-160 does through MCDPSKWaveform (detectSync -> setFrequencyOffset -> process -> getSoftBits ->
    decodeSoft): every frame's sync result, soft-bit count, ok flag, iteration count and bytes identical to the oracle on the
    identical channel outputs (LLR words are bit-identical on this path, so the decoder's results are too)."""
    import torch
    from projectultra_b200 import capi, linksim
    nc, chan, snrs = case
    ctx = capi.Context(0)
    cfg = capi.mcdpsk_config(nc, 2)
    sim = linksim.LinkSim(ctx, cfg, chan, payload_bytes=20, pool=3, code_rate=capi.R1_4, layout="chirp")
    nsym = 9 + -(-648 // (2 * nc))
    assert sim.L == 57600 + nsym * 512
    trials = 3
    si = np.repeat(np.arange(len(snrs)), trials)
    tr = np.tile(np.arange(trials), len(snrs))
    batch = sim.make_batch(list(snrs), si, tr)
    counters = torch.zeros((len(snrs), 6), dtype=torch.int64, device="cuda")
    rx, info, ok, iters = sim.run_batch(batch, counters, keep=True)
    torch.cuda.synchronize()
    rx_h = rx.cpu().numpy()
    n_llr, sync, cfo = sim.last_n_llr.cpu().numpy(), sim.last_sync.cpu().numpy(), sim.last_cfo.cpu().numpy()
    ok_h, info_h, it_h = ok.cpu().numpy(), info.cpu().numpy(), iters.cpu().numpy()
    decoded = 0
    for b in range(len(rx_h)):
        ol, oi, of, oa = O.mcdpsk_chirp_receive(nc, rx_h[b])
        assert (sync[b] == oi).all() and int(n_llr[b]) == min(len(ol), 648), (b, sync[b], oi, n_llr[b], len(ol))
        assert np.float32(cfo[b]).view(np.uint32) == np.float32(oa).view(np.uint32), (b, cfo[b], oa)
        if len(ol) >= 648:
            ci, cok, cit = O.ldpc_decode_batch(R.R1_4, ol[None, :648].copy())
            assert ok_h[b] == cok[0] and it_h[b] == cit[0] and (info_h[b] == ci[0]).all(), b
            decoded += int(cok[0])
        else:
            assert ok_h[b] == 0
    c = counters.cpu().numpy()
    # frame-by-frame parity is the point; over the fading channel the 5 Hz false-positive rule and deep fades cost frames
    assert c[:, 0].tolist() == [trials] * len(snrs) and decoded >= (trials if chan == "awgn" else 1)
    assert chan != "awgn" or c[-1, 1] == 0
    del ctx
