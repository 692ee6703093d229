"""Test helper: MC-DPSK frames behind the dual chirp, [lead silence][up chirp][gap][down chirp][gap][training][ref][data][tail],
optionally shifted in frequency (analytic signal x e^{j 2 pi df t}) and with white noise at `snr` dB on the body's mean power."""
import numpy as np


def freq_shift(x, df, fs=48000.0):
    from scipy.signal import hilbert
    a = hilbert(np.asarray(x, np.float64))
    return np.real(a * np.exp(2j * np.pi * df * np.arange(len(x)) / fs)).astype(np.float32)


def mcdpsk_chirp_frame(cfg, rng, snr, lead, cfo, total=None, tail=300, body_symbols=None):
    from projectultra_b200 import capi
    body = capi.mcdpsk_tx(cfg, capi.ldpc_encode(capi.R1_4, rng.integers(0, 256, 20, dtype=np.uint8)))
    if body_symbols is not None:
        body = body[:body_symbols * cfg.samples_per_symbol]
    w = np.concatenate([np.zeros(lead, np.float32), capi.chirp_generate(48000.0, 0.0), body, np.zeros(tail, np.float32)])
    if total is not None:
        w = np.concatenate([w, np.zeros(total - len(w), np.float32)]) if total >= len(w) else w[:total]
    if cfo:
        w = freq_shift(w, cfo)
    p = float(np.mean(body.astype(np.float64) ** 2))
    return (w + rng.normal(0.0, np.sqrt(p / 10 ** (snr / 10)), len(w))).astype(np.float32)
