"""Channel simulator: the counter-based RNG specification (Philox4x32-10 known answers, Gaussian quality), the CUDA
kernels bit-identical to the oracle's CPU twin, and the statistics of the Watterson model against the unmodified
reference (the reference's own random stream is implementation-defined and deliberately not reproduced)."""
import numpy as np
import pytest

import channelapi as CH
import oracleapi as O
import refapi as R


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert CH.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert CH.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert CH.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_gaussian_quality():
    z = CH.noise_normals(0x1234567, 200000).astype(np.float64)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.01
    assert abs((z ** 3).mean()) < 0.03 and abs((z ** 4).mean() - 3) < 0.06
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 0.01
    from scipy import stats
    assert stats.kstest(z[:50000], "norm").pvalue > 1e-3
    f = CH.fading_normals(99, 50000).astype(np.float64)
    assert np.abs(np.corrcoef(f.T) - np.eye(4)).max() < 0.02 and np.abs(f.var(axis=0) - 1).max() < 0.03
    # different seeds / streams are unrelated
    assert abs(np.corrcoef(CH.noise_normals(1, 20000), CH.noise_normals(2, 20000))[0, 1]) < 0.03


def presets():
    from projectultra_b200 import linksim
    return {n: linksim.channel_preset(n) for n in ("awgn", "good", "moderate", "poor", "flutter")}


@pytest.mark.gpu
def test_kernels_match_cpu_twin_bitwise():
    import torch
    from projectultra_b200 import capi, linksim
    ctx = capi.Context(0)
    rng = np.random.default_rng(5)
    for L in (7332, 5600, 1000, 33):
        pool = (rng.standard_normal((3, L)) * 0.3).astype(np.float32)
        for name, ch in presets().items():
            B = 6
            idx = rng.integers(0, 3, B).astype(np.uint32)
            std = rng.uniform(0.01, 0.5, B).astype(np.float32)
            seed = rng.integers(0, 2 ** 63, B, dtype=np.uint64)
            want = np.stack([CH.channel_apply(ch, pool[idx[b]], std[b], seed[b]) for b in range(B)])
            host = linksim.channel_apply(ctx, ch, pool, idx, std, seed)
            assert (host.view(np.uint32) == want.view(np.uint32)).all(), (name, L, "host path")
            dev = linksim.channel_apply(ctx, ch, torch.from_numpy(pool).cuda(), torch.from_numpy(idx.view(np.int32)).cuda(),
                                        torch.from_numpy(std).cuda(), torch.from_numpy(seed.view(np.int64)).cuda())
            torch.cuda.synchronize()
            assert (dev.cpu().numpy().view(np.uint32) == want.view(np.uint32)).all(), (name, L, "device path")
    # noise off / fading only, and a delay longer than the frame
    ch = presets()["poor"]
    ch.noise_enabled = 0
    x = (rng.standard_normal((1, 64)) * 0.3).astype(np.float32)
    got = linksim.channel_apply(ctx, ch, x, np.zeros(1, np.uint32), np.ones(1, np.float32), np.array([7], np.uint64))
    assert (got[0].view(np.uint32) == CH.channel_apply(ch, x[0], 1.0, 7).view(np.uint32)).all()
    del ctx


@pytest.mark.gpu
@pytest.mark.ref
def test_statistics_match_reference_watterson():
    """Same model, different random streams: compare what the model defines -- tap delay d+1, noise power from the
    input rms, the non-stationary fading envelope that starts at 1 -- between the CUDA kernel and the reference."""
    from projectultra_b200 import capi, linksim
    ctx = capi.Context(0)
    L = 24000
    t = np.arange(L)
    tone = (0.5 * np.sin(2 * np.pi * 1500 * t / 48000)).astype(np.float32)
    # (a) noise power: AWGN at 10 dB on a tone
    ch = presets()["awgn"]
    std = linksim.channel_noise_std(tone, 10.0, 0)
    B = 64
    got = linksim.channel_apply(ctx, ch, tone[None, :], np.zeros(B, np.uint32), np.full(B, std, np.float32),
                                np.arange(B, dtype=np.uint64))
    n_gpu = (got - tone[None, :]).astype(np.float64)
    ref = np.stack([R.watterson(tone, 10.0, 0.0, 0.0, 1.0, 0.0, fading=False, multipath=False, seed=100 + i) for i in range(8)])
    n_ref = (ref - tone[None, :]).astype(np.float64)
    assert abs(n_gpu.var() / n_ref.var() - 1) < 0.02
    assert abs(10 * np.log10((tone.astype(np.float64) ** 2).mean() / n_gpu.var()) - 10.0) < 0.1
    # (b) impulse response without noise/fading: second tap lands d+1 samples later with gain g2
    imp = np.zeros(400, np.float32)
    imp[10] = 1.0
    for name in ("good", "moderate", "poor"):
        c = presets()[name]
        c.fading_enabled = 0
        c.noise_enabled = 0
        g = linksim.channel_apply(ctx, c, imp[None, :], np.zeros(1, np.uint32), np.zeros(1, np.float32), np.zeros(1, np.uint64))[0]
        r = R.watterson(imp, 30.0, c.delay_spread_ms, c.doppler_spread_hz, fading=False, multipath=True, noise=False)
        assert (g.view(np.uint32) == r.view(np.uint32)).all(), name
        assert np.nonzero(g)[0].tolist() == [10, 10 + int(c.delay_spread_ms * 48) + 1]
    # (c) fading envelope statistics (noise off): mean |h1| trajectory from (1,0), ensemble over seeds
    for name, tol in (("flutter", 0.12), ("poor", 0.08)):
        c = presets()[name]
        c.noise_enabled = 0
        c.multipath_enabled = 0
        ones = np.ones(L, np.float32)
        B = 96
        g = linksim.channel_apply(ctx, c, ones[None, :], np.zeros(B, np.uint32), np.zeros(B, np.float32),
                                  np.arange(1000, 1000 + B, dtype=np.uint64)).astype(np.float64)
        r = np.stack([R.watterson(ones, 30.0, c.delay_spread_ms, c.doppler_spread_hz, fading=True, multipath=False,
                                  noise=False, seed=500 + i) for i in range(B)]).astype(np.float64)
        for sl in (slice(0, 200), slice(4000, 6000), slice(20000, 24000)):
            mg, mr = np.sqrt((g[:, sl] ** 2).mean()), np.sqrt((r[:, sl] ** 2).mean())
            assert abs(mg / mr - 1) < tol, (name, sl, mg, mr)
        assert abs(g[:, :5].mean() - 1.0) < 0.05     # starts at (1, 0), hf_channel.hpp:91-92
    del ctx


@pytest.mark.ref
def test_cfo_injector_twin_matches_reference_bitwise():
    """WattersonChannel::applyCFO (hf_channel.hpp:173-232) is deterministic: the oracle's restatement equals the compiled reference."""
    rng = np.random.default_rng(8)
    for L in (7332, 300, 200, 24000):
        x = (rng.standard_normal(L) * 0.3).astype(np.float32)
        for cfo in (30.0, -12.5, 50.0, 0.0005, 0.0):
            a, b = R.watterson_cfo(x, cfo), CH.channel_apply_cfo(x, cfo)
            assert (a.view(np.uint32) == b.view(np.uint32)).all(), (L, cfo)


@pytest.mark.gpu
def test_cfo_injector_kernel_matches_twin_bitwise():
    import torch
    from projectultra_b200 import capi, linksim
    ctx = capi.Context(0)
    rng = np.random.default_rng(9)
    for L in (7332, 300, 200, 7001):
        x = (rng.standard_normal((9, L)) * 0.3).astype(np.float32)
        cfo = np.array([30.0, -12.5, 50.0, 0.0005, 0.0, 7.0, -50.0, 1.0, 33.3], np.float32)
        want = np.stack([CH.channel_apply_cfo(x[b], float(cfo[b])) for b in range(9)])
        host = linksim.channel_apply_cfo(ctx, x.copy(), cfo)
        assert (host.view(np.uint32) == want.view(np.uint32)).all(), (L, "host")
        dev = linksim.channel_apply_cfo(ctx, torch.from_numpy(x).cuda(), torch.from_numpy(cfo).cuda())
        torch.cuda.synchronize()
        assert (dev.cpu().numpy().view(np.uint32) == want.view(np.uint32)).all(), (L, "device")
    del ctx
