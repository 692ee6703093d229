"""FER-curve parity at Monte-Carlo scale (BASELINE.json: "FER/BER curves must fall inside the Monte Carlo confidence interval at every SNR
point"): >= 1e4 frames per SNR point through the C++ sweep driver on the GPU (pu_linksim_run) against the UNMODIFIED reference
(oracle/_ref: sim::WattersonChannel with its own mt19937 stream -> processPresynced / demodulateSoft -> decodeSoft) on the box's host cores.
The two sides draw different noise (the GPU channel is the counter-based simulator), so the comparison is statistical: at every point the
two-proportion z statistic of the frame error rates must stay below 4 (two-sided p = 6e-5 per point) -- the intervals printed are
Wilson 95 %.  Configs 3 (M3 1024-FFT NVIS 32QAM R3/4, Watterson good) and 4 (single-carrier DQPSK R1/4, Watterson poor), plus the
headline mode over AWGN."""
import math
import multiprocessing as mp
import os

import numpy as np
import pytest

import refapi as R

pytestmark = [pytest.mark.gpu, pytest.mark.ref]
_W = {}


def _ref_worker(args):
    """Frames lo..hi of one SNR point through the reference: payload -> encode -> modulate -> Watterson -> demodulate -> decodeSoft."""
    kind, snr, lo, hi = args
    errs = 0
    rng = np.random.default_rng(1_000_003 * int(round(snr * 10 + 500)) + lo)
    for t in range(lo, hi):
        w = _W["tx"][t % len(_W["tx"])]
        payload = _W["payloads"][t % len(_W["tx"])]
        ch = _W["chan"]
        rx = R.watterson(w, snr, ch[0], ch[1], fading=ch[2], multipath=ch[2], seed=int(rng.integers(1, 2 ** 31)))
        if kind == "ofdm":
            llr, _, _ = R.ofdm_presynced(_W["cfg"], rx, 2, 1, 0.0, 0.0)
        else:
            llr = R.dpsk_demod_soft_ex(1, 384, rx, 39 * 384, ref_mode=1)
        if len(llr) < 648:
            errs += 1
            continue
        info, ok, _ = R.ldpc_decode_soft(_W["rate"], llr[:648])
        if not ok or not (info[:len(payload)] == payload).all():
            errs += 1
    return errs


def reference_curve(kind, cfg, rate, nbytes, chan, snrs, frames, peak):
    rng = np.random.default_rng(77)
    payloads = rng.integers(0, 256, (16, nbytes), dtype=np.uint8)
    tx = []
    for p in payloads:
        coded = R.ldpc_encode(rate, p)
        w = R.ofdm_tx(cfg, coded, 0) if kind == "ofdm" else R.dpsk_tx(1, 384, coded, 0)
        if peak:
            w = (w * (np.float32(peak) / np.abs(w).max())).astype(np.float32)
        tx.append(w)
    _W.update(tx=tx, payloads=payloads, cfg=cfg, rate=rate, chan=chan)
    cores = max(1, len(os.sched_getaffinity(0)))
    per = max(1, frames // (cores * 2))
    jobs = [(kind, s, lo, min(frames, lo + per)) for s in snrs for lo in range(0, frames, per)]
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_ref_worker, jobs)
    errs = {s: 0 for s in snrs}
    for (k, s, lo, hi), e in zip(jobs, res):
        errs[s] += e
    return [errs[s] for s in snrs]


def check_curves(name, snrs, frames_ref, errs_ref, counters):
    from projectultra_b200 import capi
    worst = 0.0
    for i, s in enumerate(snrs):
        n2, e2 = int(counters[i, 0]), int(counters[i, 1])
        n1, e1 = frames_ref, errs_ref[i]
        p1, p2 = e1 / n1, e2 / n2
        pool = (e1 + e2) / (n1 + n2)
        se = math.sqrt(max(pool * (1 - pool), 1e-12) * (1 / n1 + 1 / n2))
        z = abs(p1 - p2) / se if (e1 + e2) > 0 and (e1 + e2) < (n1 + n2) else 0.0
        lo1, hi1 = capi.wilson_interval(e1, n1)
        lo2, hi2 = capi.wilson_interval(e2, n2)
        print("%s %6.1f dB  reference FER %.4f [%.4f, %.4f] (%d frames)   GPU FER %.4f [%.4f, %.4f] (%d frames)   z = %.2f"
              % (name, s, p1, lo1, hi1, n1, p2, lo2, hi2, n2, z))
        worst = max(worst, z)
    assert worst < 4.0, (name, worst)


@pytest.fixture(scope="module")
def ctx():
    from projectultra_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def test_config3_m3_32qam_r34_watterson_good(ctx):
    from projectultra_b200 import capi
    frames = 10000
    snrs = [18.0, 24.0, 30.0, 36.0, 42.0]      # the reference's own curve is still near 1 at 26 dB on this channel
    cfg = capi.ModemConfig(48000, 1500, 1024, 59, 1, 0, 4, 1, capi.QAM32, capi.R3_4, 40.0, 0.0)
    mode = capi.sweep_mode(capi.WF_OFDM, cfg, capi.R3_4, 60, "good", snrs[0], snrs[1] - snrs[0], len(snrs))
    counters, st = capi.Sweep([mode], trials_per_point=frames, block_trials=2500, pool=16).run(ctx)
    assert (counters[:, 0] == frames).all()
    rcfg = R.config_m3(R.QAM32, R.R3_4)
    errs = reference_curve("ofdm", rcfg, R.R3_4, 60, (0.5, 0.1, True), snrs, frames, None)
    check_curves("config 3", snrs, frames, errs, counters)


def test_config4_dqpsk_r14_watterson_poor(ctx):
    from projectultra_b200 import capi
    frames = 10000
    snrs = [-26.0, -23.0, -20.0, -17.0, -14.0]  # 125 baud in 24 kHz of noise: 23 dB of processing gain, error free from -11 dB up
    mode = capi.sweep_mode(capi.WF_DPSK, capi.dpsk_config(1, 384), capi.R1_4, 20, "poor", snrs[0], snrs[1] - snrs[0], len(snrs), peak=0.5)
    counters, st = capi.Sweep([mode], trials_per_point=frames, block_trials=1250, pool=16).run(ctx)
    assert (counters[:, 0] == frames).all()
    errs = reference_curve("dpsk", None, R.R1_4, 20, (2.0, 1.0, True), snrs, frames, 0.5)
    check_curves("config 4", snrs, frames, errs, counters)


def test_headline_m1_dqpsk_r12_awgn_fast_precision(ctx):
    """The bench workload itself, in the arithmetic the bench runs (PU_PRECISION_FAST).  WattersonChannel without fading / multipath is
    the reference's AWGN channel with the rms SNR convention; the GPU side uses the tools' mean-power convention (same number)."""
    from projectultra_b200 import capi
    frames = 20000
    snrs = [-1.0, 0.0, 1.0, 2.0, 3.0]
    cfg = capi.ModemConfig(48000, 1500, 512, 30, 1, 4, 2, 0, capi.DQPSK, capi.R1_2, 40.0, 0.0)
    mode = capi.sweep_mode(capi.WF_OFDM, cfg, capi.R1_2, 40, "awgn", snrs[0], 1.0, len(snrs), precision="fast")
    counters, st = capi.Sweep([mode], trials_per_point=frames, block_trials=4000, pool=16).run(ctx)
    rcfg = R.config_m1(R.DQPSK, R.R1_2)
    errs = reference_curve("ofdm", rcfg, R.R1_2, 40, (0.0, 0.0, False), snrs, frames, None)
    check_curves("headline", snrs, frames, errs, counters)
