"""GPU parity tests of PU_PRECISION_FAST (csrc/ofdm_fast512.cu: FMA-contracted butterflies) at BASELINE.json's own bar:
LLRs within 1e-4 relative, saturated LLRs exactly the reference's +-10, hard decisions after LDPC bit-exact on every frame
the reference decodes with margin.  References: the plain-C oracle (pinned bit-exact to the compiled reference) on small
batches, the bit-exact kernel (ofdm_diff512.cu, itself pinned to the oracle by test_ofdm_gpu.py / test_ofdm_bitexact_gpu.py)
on the bench-sized batch.

Two kinds of LLR can legitimately differ by more than the relative bound although the underlying soft value moved by < 1e-4:
  * clip-floor flips: soft_demap::clipLLR (soft_demap.hpp:22-29) maps every |llr| < 0.5 to +-0.5, so a soft value within
    rounding distance of 0 comes out as +0.5 in one arithmetic and -0.5 in the other;
  * gate flips: the weak-signal gate (|sym||prev| < 1e-6, :178,199,224) zeroes a carrier in one arithmetic only.
They are counted separately and bounded (a few per million LLRs); everything else must be inside the bound."""
import numpy as np
import pytest

import oracleapi as O
import refapi as R
from golden.make_golden import awgn

pytestmark = pytest.mark.gpu
LLR_RTOL = 1e-4     # BASELINE.json north_star: "LLRs must fall within a stated relative tolerance (1e-4 on fp32)"


@pytest.fixture(scope="module")
def ctx():
    from projectultra_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def to_capi_cfg(cfg):
    from projectultra_b200 import capi
    return capi.ModemConfig.from_buffer_copy(bytes(cfg))


def classify(got, ref):
    """-> (n_out_of_tolerance_unexplained, n_clip_floor_flips, n_gate_flips, fraction_bit_identical)"""
    got = np.asarray(got, np.float32).ravel()
    ref = np.asarray(ref, np.float32).ravel()
    same = got.view(np.uint32) == ref.view(np.uint32)
    bad = np.abs(got.astype(np.float64) - ref) > LLR_RTOL * np.maximum(np.abs(ref), 0.5)
    floor_flip = bad & (np.abs(ref) == 0.5) & (got == -ref)
    gate_flip = bad & ((ref == 0) != (got == 0))
    return int((bad & ~floor_flip & ~gate_flip).sum()), int(floor_flip.sum()), int(gate_flip.sum()), float(same.mean())


def truth_llrs(cfg, bins_idx, frames, sym_len, training=2):
    """The reference's LLR laws (soft_demap.hpp:173-237 in closed form: |d| = |sym||prev|, so sp cos(phase) = Re d, ...) evaluated
    in float64 from a float64 DFT of the mixed samples: what both fp32 arithmetics approximate.  The mixer table is the
    reference's (float phase accumulation, filters.cpp:228-238)."""
    nfft = int(cfg.fft_size)
    cp = sym_len - nfft - int(cfg.symbol_guard)
    inc = np.float32(2.0 * np.pi * float(cfg.center_freq) / np.float32(cfg.sample_rate))
    ph = np.zeros(frames.shape[1], np.float32)
    p = np.float32(0)
    for i in range(len(ph)):
        ph[i] = p
        p = np.float32(p + inc)
        if p > 2.0 * np.pi:
            p = np.float32(np.float64(p) - 2.0 * np.pi)
    osc = np.cos(ph).astype(np.float64) - 1j * np.sin(ph).astype(np.float64)           # conj(osc)
    nc = int(cfg.num_carriers)
    n = np.arange(nc, dtype=np.float64)
    zc = np.exp(1j * (-np.pi * n * (n + 1) / nc).astype(np.float32).astype(np.float64))
    margin = 1.1 if cfg.modulation == R.D8PSK else 1.0
    n_sym = frames.shape[1] // sym_len
    x = frames.astype(np.float64) * osc[None, :]
    win = np.stack([x[:, s * sym_len + cp: s * sym_len + cp + nfft] for s in range(training - 1, n_sym)], axis=1)
    bins = np.fft.fft(win, axis=2)[:, :, bins_idx]                                   # [frame][symbol][carrier]
    h = bins[:, 0] / zc[None, :len(bins_idx)]
    nv = (np.clip(0.1 / np.abs(h) ** 2, 1e-6, 100.0) * margin)[:, None, :]
    eq = bins[:, 1:] / h[:, None, :]
    prev = np.concatenate([np.ones_like(eq[:, :1]), eq[:, :-1]], axis=1)
    d = eq * np.conj(prev)
    sp, phase = np.abs(d), np.angle(d)
    if cfg.modulation == R.DBPSK:
        l = (2 * sp * np.cos(phase) / nv)[..., None]
    elif cfg.modulation == R.DQPSK:
        l = np.stack([2 * sp / nv * np.sin(phase + np.pi / 4), 2 * sp / nv * np.cos(2 * phase)], axis=-1)
    else:
        l = np.stack([sp / nv * np.sin(phase), sp / nv * np.sin(2 * phase), sp / nv * np.sin(4 * phase)], axis=-1)
    return l.reshape(len(frames), -1)            # unclipped soft values


def frames_for(cfg, rate, snrs, per_snr, seed0):
    out = []
    for i, snr in enumerate(snrs):
        for j in range(per_snr):
            rng = np.random.default_rng(seed0 + 1000 * i + j)
            data = rng.integers(0, 256, 40, dtype=np.uint8)
            out.append(awgn(O.ofdm_tx(cfg, O.ldpc_encode(rate, data), 0), snr, rng))
    return np.stack(out)


@pytest.mark.parametrize("mod", [R.DBPSK, R.DQPSK, R.D8PSK])
@pytest.mark.parametrize("deint", [False, True])
def test_fast_llrs_against_oracle(ctx, mod, deint):
    from projectultra_b200 import capi
    cfg = R.config_m1(mod, R.R1_2)
    dem = capi.OfdmDemodulator(ctx, to_capi_cfg(cfg))
    dem.set_precision("fast")
    assert dem.precision == "fast"
    frames = frames_for(cfg, R.R1_2, (-4.0, 0.0, 3.0, 6.0, 10.0, 20.0), 8 + (mod == R.DQPSK) * 8, 7000 + 10 * mod)
    n_llr = dem.n_llr(frames.shape[1])
    ref, counts = O.ofdm_presynced_batch(cfg, frames, n_llr)
    assert (counts == n_llr).all()
    if deint:      # fused ChannelInterleaver::deinterleave over the first codeword, truncated output
        bps = {R.DBPSK: 1, R.DQPSK: 2, R.D8PSK: 3}[mod] * dem.n_data
        dem.set_deinterleave(bps, 648)
        llr, _, _ = dem.presynced_batch(frames, llr_stride=648)
        assert dem.last_kernel == "ofdm_fast512_kernel"
        want = np.stack([O.channel_interleave(bps, r[:648].copy(), inverse=True) for r in ref])
        bad, floor, gate, same = classify(llr, want)
        assert bad <= max(3, int(1e-3 * want.size)) and floor + gate <= 2, (bad, floor, gate, same)
        return
    llr, snr, fc = dem.presynced_batch(frames)
    assert dem.last_kernel == "ofdm_fast512_kernel"
    bad, floor, gate, same = classify(llr, ref)
    d_max = float(np.abs(llr.astype(np.float64) - ref).max())
    print("mod %d: %d LLRs: outside 1e-4 rel %d, clip-floor flips %d, gate flips %d, bit-identical %.5f, max |fast-ref| %.2e"
          % (mod, ref.size, bad, floor, gate, same, d_max))
    # 1e-4 relative (floor 0.5) up to the few LLRs whose soft value sits next to a zero of its law: with |h|^2 ~ 1e3 (unnormalised FFT)
    # the laws' prefactor 2 sp / nv is ~1e3..1e4, every LLR that is not clipped to +-10 is such a one, and ANY reordering of the
    # FFT's roundings (relative 3e-7) moves it by prefactor * 3e-7 ~ 1e-3; test_fast_hard_decisions... bounds them against float64
    assert bad <= max(3, int(1e-3 * ref.size)), (bad, ref.size)
    assert floor + gate <= max(2, int(2e-5 * ref.size)), (floor, gate, ref.size)
    assert same > 0.9, same                                    # saturated LLRs are exactly +-10 in both
    # SNR report and CFO as the exact kernels give them (|h| is invariant under the lane sign of the fast kernel)
    dem.set_precision("exact")
    _, snr_x, fc_x = dem.presynced_batch(frames)
    assert np.allclose(snr, snr_x, rtol=1e-5, atol=1e-4) and (fc == 0).all() and (fc_x == 0).all()
    # truncated output (first codeword only), odd batch size, single frame
    dem.set_precision("fast")
    for nb in (1, 3, len(frames)):
        l648, _, _ = dem.presynced_batch(frames[:nb], llr_stride=648)
        assert (l648.view(np.uint32) == llr[:nb, :648].view(np.uint32)).all(), nb


def test_fast_hard_decisions_match_on_frames_decoded_with_margin(ctx):
    """53 248-frame bench batch (13 SNR points x 4 096 frames, AWGN): fast against the bit-exact kernel on the same channel outputs."""
    import torch
    from projectultra_b200 import capi, linksim
    cfg = capi.ModemConfig(48000, 1500, 512, 30, 1, 4, 2, 0, capi.DQPSK, capi.R1_2, 40.0, 0.0)
    sim = linksim.LinkSim(ctx, cfg, "awgn", payload_bytes=40, pool=64, code_rate=capi.R1_2)
    snrs = [float(s) for s in range(-4, 9)]
    fpp = 4096
    si = np.tile(np.arange(len(snrs), dtype=np.int64), fpp)
    tr = np.repeat(np.arange(fpp, dtype=np.int64), len(snrs))
    batch = sim.make_batch(snrs, si, tr)
    rx = linksim.channel_apply(ctx, sim.ch, sim.tx_pool, batch["tx_index"], batch["noise_std"], batch["seed"])
    out = {}
    for mode in ("exact", "fast"):
        sim.ofdm.set_precision(mode)
        llr = sim.ofdm.presynced_batch(rx, 2, llr_stride=648, want_aux=False)[0]
        assert sim.ofdm.last_kernel == ("ofdm_fast512_kernel" if mode == "fast" else "ofdm_diff512_kernel")
        info, ok, it = sim.ldpc.decode_batch(llr)
        torch.cuda.synchronize()
        out[mode] = (llr.cpu().numpy(), info.cpu().numpy(), ok.cpu().numpy(), it.cpu().numpy())
    lx, ix, okx, itx = out["exact"]
    lf, inf_, okf, itf = out["fast"]
    bad, floor, gate, same = classify(lf, lx)
    ok_abs = np.abs(lf.astype(np.float64) - lx) <= 1e-3
    print("fast vs exact over %d LLRs: outside 1e-4 rel %d (%.2e), of which beyond 1e-3 abs %d; clip-floor flips %d, gate flips %d, bit-identical %.6f"
          % (lx.size, bad, bad / lx.size, int((~ok_abs).sum()) - floor - gate, floor, gate, same))
    assert bad <= 3e-4 * lx.size
    assert np.abs(lf.astype(np.float64) - lx).max() <= 2e-2 or floor + gate > 0
    assert floor + gate <= 2e-5 * lx.size
    # distance of both fp32 arithmetics from the float64 evaluation of the same laws, over the LLRs neither clip touches: the fast
    # kernel is as close to the exact soft values as the reference's own arithmetic is
    nt = 8192
    tru = truth_llrs(R.config_m1(R.DQPSK, R.R1_2), sim.ofdm.carrier_bins()[:sim.ofdm.n_data], rx[:nt].cpu().numpy(),
                     sim.ofdm.symbol_samples)[:, :648]
    free = (np.abs(tru) > 0.6) & (np.abs(tru) < 9.9)
    e_ref = np.abs(lx[:nt].astype(np.float64) - tru)[free]
    e_fast = np.abs(lf[:nt].astype(np.float64) - tru)[free]
    rms = lambda e: float(np.sqrt((e ** 2).mean()))
    print("against float64 over %d unclipped LLRs of %d: reference rms %.3e max %.3e | fast rms %.3e max %.3e"
          % (free.sum(), tru.size, rms(e_ref), e_ref.max(), rms(e_fast), e_fast.max()))
    assert free.sum() > 500
    assert rms(e_fast) <= 1.1 * rms(e_ref) and e_fast.max() <= 1.5 * e_ref.max()
    clipped = np.abs(tru) > 10.5
    assert (lf[:nt][clipped] == np.sign(tru[clipped]) * 10).all() and (lx[:nt][clipped] == np.sign(tru[clipped]) * 10).all()
    # frames the reference decodes with margin: converged at least 10 iterations before the limit
    margin = (okx == 1) & (itx <= 40)
    assert margin.sum() > 0.4 * len(okx)
    assert (okf[margin] == 1).all() and (inf_[margin] == ix[margin]).all()
    # everything else: the frame-error counters agree to within the frames whose LLRs differ at all
    differs = (lf.view(np.uint32) != lx.view(np.uint32)).any(axis=1)
    flips = int((okf != okx).sum())
    print("frames with any differing LLR word: %d of %d; decode-verdict flips: %d" % (differs.sum(), len(okx), flips))
    assert flips <= max(3, int(2e-4 * len(okx)))
    assert ((okf == okx) | differs).all()
    # determinism
    llr2 = sim.ofdm.presynced_batch(rx, 2, llr_stride=648, want_aux=False)[0]
    torch.cuda.synchronize()
    assert (llr2.cpu().numpy().view(np.uint32) == lf.view(np.uint32)).all()


GENERAL_CASES = [("m1", R.QPSK, R.R1_2, 40, (4.0, 10.0, 18.0), None), ("m1", R.QAM16, R.R1_2, 40, (8.0, 14.0, 22.0), None),
                 ("m1", R.BPSK, R.R1_2, 40, (0.0, 6.0, 12.0), None), ("m1", R.QAM64, R.R3_4, 60, (18.0, 24.0, 30.0), None),
                 ("m3", R.QAM32, R.R3_4, 60, (12.0, 18.0, 26.0), None), ("m3", R.QAM16, R.R3_4, 60, (10.0, 16.0, 24.0), None),
                 ("m1", R.DQPSK, R.R1_2, 40, (2.0, 8.0, 16.0), (6.5, -1.3)), ("m1", R.QAM16, R.R1_2, 40, (10.0, 16.0, 24.0), (-12.0, 0.4)),
                 ("m3", R.QAM32, R.R3_4, 60, (14.0, 20.0, 28.0), (3.0, 2.0))]


@pytest.mark.parametrize("case", range(len(GENERAL_CASES)))
def test_fast_general_kernel_against_exact_kernel_and_oracle(ctx, case):
    """PU_PRECISION_FAST of the general warp kernel (pilots, coherent QAM, CFO rotator: FMA butterflies + closed-form rotator) against
    the exact kernel on the same frames, and against the oracle on a subset: LLRs within 1e-4 * max(|ref|, 0.5) up to the few that sit
    next to a zero of their law or ride on the tracked CFO's last bits, LDPC verdicts and bytes identical on frames decoded with margin,
    the tracked CFO and SNR reports equal to ~1e-4."""
    from projectultra_b200 import capi
    preset, mod, rate, nbytes, snrs, cfo = GENERAL_CASES[case]
    cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
    if cfo:
        cfg.tx_cfo_hz = cfo[0]
    dem = capi.OfdmDemodulator(ctx, to_capi_cfg(cfg))
    frames = []
    for i, snr in enumerate(snrs):
        for j in range(12):
            rng = np.random.default_rng(5000 + 100 * case + 20 * i + j)
            data = rng.integers(0, 256, nbytes, dtype=np.uint8)
            frames.append(awgn(O.ofdm_tx(cfg, O.ldpc_encode(rate, data), 0), snr, rng))
    frames = np.stack(frames)
    B = len(frames)
    cf = np.full(B, cfo[0], np.float32) if cfo else None
    ph = np.full(B, cfo[1], np.float32) if cfo else None
    lx, snr_x, fc_x = dem.presynced_batch(frames, 2, cf, ph)
    assert dem.last_kernel == "ofdm_presynced_warp_kernel"
    dem.set_precision("fast")
    lf, snr_f, fc_f = dem.presynced_batch(frames, 2, cf, ph)
    assert dem.last_kernel == "ofdm_presynced_warp_fast_kernel"
    bad, floor, gate, same = classify(lf, lx)
    d = np.abs(lf.astype(np.float64) - lx)
    print("case %d (%s mod %d cfo %s): %d LLRs, outside 1e-4 rel %d (%.2e), clip-floor flips %d, gate flips %d, bit-identical %.4f, max |d| %.2e, "
          "p99.9 |d| %.2e; tracked CFO max |d| %.2e Hz, SNR max |d| %.2e dB"
          % (case, preset, mod, cfo, lx.size, bad, bad / lx.size, floor, gate, same, d.max(), np.quantile(d, 0.999),
             np.abs(fc_f - fc_x).max(), np.abs(snr_f - snr_x).max()))
    assert bad <= 5e-3 * lx.size, (bad, lx.size)
    assert np.quantile(d, 0.999) <= 2e-3 and d.max() <= 0.05 or floor + gate > 0
    assert np.abs(fc_f - fc_x).max() <= 2e-3 and np.abs(snr_f - snr_x).max() <= 1e-2
    dec = capi.LdpcDecoder(ctx, rate)
    ix, okx, itx = dec.decode_batch(lx[:, :648].copy())
    i_f, okf, itf = dec.decode_batch(lf[:, :648].copy())
    margin = (okx == 1) & (itx <= 40)
    assert (okf[margin] == 1).all() and (i_f[margin] == ix[margin]).all()
    assert (okf != okx).sum() <= 1
    # the oracle on the first frame of every SNR point
    rcfg = cfg
    for b in range(0, B, 12):
        want, _, _ = O.ofdm_presynced(rcfg, frames[b], 2, 2 if cfo else 1, cfo[0] if cfo else 0.0, cfo[1] if cfo else 0.0)
        bo, fo, go, _ = classify(lf[b][:len(want)], want)
        assert bo <= 5e-3 * len(want) + 2, (b, bo)


def test_fast_falls_back_to_exact_kernels_outside_its_coverage(ctx):
    from projectultra_b200 import capi
    cfg = R.config_m3(R.DQPSK, R.R3_4)          # 1024-FFT differential no-pilot at zero CFO: ofdm_diff_kernel has no fast form
    dem = capi.OfdmDemodulator(ctx, to_capi_cfg(cfg))
    frames = frames_for(cfg, R.R3_4, (6.0, 15.0), 2, 9100)
    a = dem.presynced_batch(frames)[0]
    dem.set_precision("fast")
    b = dem.presynced_batch(frames)[0]
    assert dem.last_kernel == "ofdm_diff_kernel"
    assert (a.view(np.uint32) == b.view(np.uint32)).all()
