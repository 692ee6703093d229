"""SURVEY §8f next-2 (chirp half): dual-chirp synchronisation of the OFDM_CHIRP waveform on the GPU -- the receive sequence
of tools/test_iwaveform.cpp:127-160 (IWaveform::detectSync -> setFrequencyOffset -> process -> getSoftBits) -- against the
plain-C oracle (oracle/pu_oracle_ofdm.c: orc_ofdm_chirp_receive, pinned to the compiled reference) and, when present, the compiled
reference.  Detection flag, chirp positions, training start: identical integers; CFO: identical bits; LLRs within 1e-4."""
import numpy as np
import pytest

import refapi as R
import oracleapi as O

pytestmark = pytest.mark.gpu


def frame(cfg, rate, nbytes, snr, seed, lead, tail, tx_cfo=0.0):
    from projectultra_b200 import capi
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, nbytes, dtype=np.uint8)
    body = O.ofdm_tx(cfg, O.ldpc_encode(rate, data), 0)                 # generateTrainingSymbols(2) + modulate()
    chirp = capi.chirp_generate(48000.0, tx_cfo)
    w = np.concatenate([np.zeros(lead, np.float32), chirp, body, np.zeros(tail, np.float32)])
    if snr is None:
        return w
    p = float(np.mean(body.astype(np.float64) ** 2))
    return (w + rng.normal(0.0, np.sqrt(p / 10 ** (snr / 10)), len(w))).astype(np.float32)


def test_chirp_generate_matches_oracle():
    from projectultra_b200 import capi
    for cfo in (0.0, 12.5, -30.0):
        a, b = capi.chirp_generate(48000.0, cfo), O.chirp_generate(48000.0, cfo)
        assert a.shape == b.shape and (a.view(np.uint32) == b.view(np.uint32)).all()


@pytest.mark.parametrize("preset,mod", [("m1", R.DQPSK), ("m1", R.QAM16), ("m3", R.DQPSK)])
def test_chirp_receive_matches_oracle(preset, mod):
    import torch
    from projectultra_b200 import capi
    rate = R.R1_2
    ctx = capi.Context(0)
    total = 57600 + 16000
    frames, cfgs = [], []
    cases = [(25.0, 0, 0.0), (12.0, 3000, 0.0), (4.0, 777, 0.0), (-4.0, 5000, 0.0), (-12.0, 100, 0.0), (20.0, 1200, 12.5), (15.0, 40, -30.0), (None, 2500, 0.0)]
    for i, (snr, lead, tx_cfo) in enumerate(cases):
        cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
        cfg.tx_cfo_hz = tx_cfo
        body_len = len(O.ofdm_tx(cfg, O.ldpc_encode(rate, np.zeros(40, np.uint8)), 0))
        frames.append(frame(cfg, rate, 40, snr, 800 + 13 * i + mod, lead, total - 57600 - body_len - lead, tx_cfo))
        cfgs.append(cfg)
    frames.append(np.random.default_rng(5).normal(0, 0.1, total).astype(np.float32))     # noise only
    frames.append(np.zeros(total, np.float32))                                            # silence
    cfgs += [cfgs[0], cfgs[0]]
    x = np.stack(frames)
    base = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
    dem = capi.OfdmDemodulator(ctx, capi.ModemConfig.from_buffer_copy(bytes(base)))
    llr, n, info, val, snr_db = dem.chirp_receive_batch(x, llr_stride=700)
    found = 0
    for b in range(len(x)):
        ol, oi, ocfo = O.ofdm_chirp_receive(base, x[b])
        assert (info[b] == oi).all(), (b, info[b], oi)
        assert np.float32(val[b, 0]).view(np.uint32) == np.float32(ocfo).view(np.uint32), (b, val[b, 0], ocfo)
        want = ol[:700]
        assert int(n[b]) == len(want), (b, n[b], len(ol))
        if len(want):
            found += 1
            bad = np.flatnonzero(~np.isclose(llr[b, :len(want)], want, rtol=1e-4, atol=1e-6))
            assert len(bad) == 0, (b, bad[:8], llr[b, bad[:8]], want[bad[:8]])
        if R.available() and b in (0, 5):
            rl, ri, rcfo = R.ofdm_chirp_receive(base, x[b])
            assert (ri == oi).all() and np.float32(rcfo).view(np.uint32) == np.float32(ocfo).view(np.uint32)
    assert found >= 5
    d = dem.chirp_receive_batch(torch.from_numpy(x).cuda(), llr_stride=700)
    torch.cuda.synchronize()
    assert (d[2].cpu().numpy() == info).all() and (d[1].cpu().numpy() == n).all()
    assert (d[0].cpu().numpy().view(np.uint32) == llr.view(np.uint32)).all()
    del ctx


def test_two_tier_search_equals_brute_force_search():
    """The two-tier coarse search (decimated ranking, exact evaluation of the leaders, csrc/chirp_sync.cu) against the brute-force kernel
    that evaluates every coarse position the reference's way (PU_CHIRP_SEARCH=exact): the exact first maximum must always lie inside
    the verified set, i.e. every output -- flags, positions, CFO bits, correlation bits, LLR bits -- is identical, over frames from
    +20 to -25 dB, with TX CFO, noise only, silence, a chirp pair cut short and a lone up chirp."""
    import os
    from projectultra_b200 import capi
    cfg = R.config_m1(R.DQPSK, R.R1_2)
    ctx = capi.Context(0)
    total = 57600 + 11000
    rng = np.random.default_rng(77)
    body = O.ofdm_tx(cfg, O.ldpc_encode(R.R1_2, np.zeros(40, np.uint8)), 0)
    frames = []
    for i in range(160):
        snr = float(rng.uniform(-25.0, 20.0))
        lead = int(rng.integers(0, total - 57600 - len(body)))
        tx_cfo = float(rng.choice([0.0, 0.0, 7.5, -18.0, 33.0]))
        frames.append(frame(cfg, R.R1_2, 40, snr, 9000 + i, lead, total - 57600 - len(body) - lead, tx_cfo))
    for i in range(24):
        frames.append(rng.normal(0, 0.05 + 0.1 * i, total).astype(np.float32))
    frames.append(np.zeros(total, np.float32))
    cut = frames[0].copy(); cut[40000:] = 0.0; frames.append(cut)                         # down chirp cut off
    lone = rng.normal(0, 0.05, total).astype(np.float32); lone[1000:25000] += capi.chirp_generate()[:24000]; frames.append(lone)
    x = np.stack(frames)
    dem = capi.OfdmDemodulator(ctx, capi.ModemConfig.from_buffer_copy(bytes(cfg)))
    os.environ.pop("PU_CHIRP_SEARCH", None)
    capi.chirp_search_stats()
    fast = dem.chirp_receive_batch(x, llr_stride=700)
    searches, rounds, fine_runs = capi.chirp_search_stats()
    os.environ["PU_CHIRP_SEARCH"] = "exact"
    try:
        slow = dem.chirp_receive_batch(x, llr_stride=700)
    finally:
        os.environ.pop("PU_CHIRP_SEARCH", None)
    # the multi-round paths (a further 16 coarse positions / a further fine run whenever the stop rule does not hold) are rare by design:
    # force them on every search and require the same outputs once more
    os.environ["PU_CHIRP_GUARD"] = "1e30"
    try:
        capi.chirp_search_stats()
        forced = dem.chirp_receive_batch(x[:48], llr_stride=700)
        s2, r2, f2 = capi.chirp_search_stats()
    finally:
        os.environ.pop("PU_CHIRP_GUARD", None)
    assert r2 >= 40 * s2 and f2 >= 5 * (s2 // 2), (s2, r2, f2)            # ~55 rounds of 16 = every coarse position; ~7 runs = the whole fine range
    for a, b in zip(forced[:4], fast[:4]):
        assert (np.asarray(a).view(np.uint32) == np.asarray(b)[:48].view(np.uint32)).all()
    llr_f, n_f, info_f, val_f = fast[:4]
    llr_s, n_s, info_s, val_s = slow[:4]
    assert (info_f == info_s).all(), np.flatnonzero((info_f != info_s).any(axis=1))
    assert (n_f == n_s).all()
    assert (np.asarray(val_f).view(np.uint32) == np.asarray(val_s).view(np.uint32)).all()
    assert (llr_f.view(np.uint32) == llr_s.view(np.uint32)).all()
    found = int((info_f[:, 0] != 0).sum())
    print("two-tier == brute force on %d frames, %d with both chirps found; %d template searches, %d coarse verification rounds of 16 positions, "
          "%d fine runs of 16 positions" % (len(x), found, searches, rounds, fine_runs))
    assert searches >= len(x) - 1 and rounds <= 2 * searches        # the ranking does the work: ~1 round per search, not ~30
    assert 60 <= found < len(x) - 10
    del ctx


def test_two_tier_search_equals_brute_force_search_on_two_path_channels():
    """Two propagation paths 24 samples apart (Watterson "good") put two comparable peaks and their interference pattern (32-sample period)
    inside the fine search range: the case that defeated a 6-sample estimate grid with a main-lobe reach factor (2 of 16 384 frames took the
    lower peak).  With estimates every 3 samples and the Bernstein bound as reach factor every output equals the brute-force search."""
    import os
    import torch
    from projectultra_b200 import capi, linksim
    ctx = capi.Context(0)
    cfg = capi.ModemConfig.from_buffer_copy(bytes(R.config_m1(R.DQPSK, R.R1_2)))
    differ = 0
    for chan, snr, B in (("good", 12.0, 3072), ("poor", 0.0, 1024)):
        sim = linksim.LinkSim(ctx, cfg, chan, payload_bytes=40, pool=16, layout="chirp", peak=0.5, precision="fast")
        pool = torch.cat([sim.tx_pool, torch.zeros(sim.tx_pool.shape[0], 2400, device=sim.tx_pool.device)], dim=1).contiguous()
        idx = (torch.arange(B, device=pool.device) % pool.shape[0]).to(torch.int32)
        std = ((pool.double() ** 2).mean(dim=1).sqrt() * 10 ** (-snr / 20)).float()[idx.long()].contiguous()
        seed = torch.arange(B, device=pool.device, dtype=torch.int64) + 4242
        rx = linksim.channel_apply(ctx, sim.ch, pool, idx, std, seed, None)
        os.environ.pop("PU_CHIRP_SEARCH", None)
        capi.chirp_search_stats()
        fast = sim.ofdm.chirp_receive_batch(rx, llr_stride=648)
        searches, rounds, runs = capi.chirp_search_stats()
        os.environ["PU_CHIRP_SEARCH"] = "exact"
        try:
            slow = sim.ofdm.chirp_receive_batch(rx, llr_stride=648)
        finally:
            os.environ.pop("PU_CHIRP_SEARCH", None)
        torch.cuda.synchronize()
        fi, si = fast[2].cpu().numpy(), slow[2].cpu().numpy()
        fv, sv = fast[3].cpu().numpy().view(np.uint32), slow[3].cpu().numpy().view(np.uint32)
        bad = np.flatnonzero((fi != si).any(axis=1) | (fv != sv).any(axis=1))
        print("%s %.0f dB: %d frames, %d detected, %d differ; %d searches, %d coarse rounds, %d fine runs" % (
            chan, snr, B, int((si[:, 0] != 0).sum()), len(bad), searches, rounds, runs))
        differ += len(bad)
        assert int((si[:, 0] != 0).sum()) > B // 2
    assert differ == 0
    del ctx
