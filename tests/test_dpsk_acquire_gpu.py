"""SURVEY §8f next-2 (DPSK half): Barker-13 acquisition of the single-carrier DPSK waveform on the GPU -- the receive sequence
of tools/test_dpsk_snr.cpp:66-73 (findPreamble on the whole frame, demodulateSoft from the returned data start) -- against the
plain-C oracle (oracle/pu_oracle_psk.c, itself pinned to the compiled reference) and, when present, the compiled reference.
Data start, estimated CFO and initial phase offset must be identical (integers / float bits), LLR words bit-identical."""
import numpy as np
import pytest

import refapi as R
import oracleapi as O

pytestmark = pytest.mark.gpu


def words(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def make(cfg, mod, snr, seed, lead, total):
    from projectultra_b200 import capi
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, 20, dtype=np.uint8)
    tx = capi.dpsk_tx(cfg, capi.ldpc_encode(capi.R1_4, data), 0)          # generatePreamble() + modulate()
    tx = (tx * (np.float32(0.5) / np.abs(tx).max())).astype(np.float32)   # tools/test_dpsk_snr.cpp:52-56
    w = np.zeros(total, np.float32)
    w[lead:lead + len(tx)] = tx
    if snr is None:
        return w
    p = float(np.mean(tx.astype(np.float64) ** 2))
    return (w + rng.normal(0.0, np.sqrt(p / 10 ** (snr / 10)), total)).astype(np.float32)


@pytest.mark.parametrize("mod", [0, 1, 2])
def test_dpsk_receive_matches_oracle(mod):
    import torch
    from projectultra_b200 import capi
    ctx = capi.Context(0)
    cfg = capi.dpsk_config(mod, 384)
    dem = capi.DpskDemodulator(ctx, cfg)
    frame = 39 * 384 + (648 // (mod + 1)) * 384
    total = frame + 1200
    cases = [(20.0, 0), (10.0, 137), (3.0, 384), (-2.0, 901), (-6.0, 5), (-10.0, 640), (-14.0, 333), (-20.0, 100), (None, 77)]
    frames = [make(cfg, mod, snr, 60 + 7 * i + mod, lead, total) for i, (snr, lead) in enumerate(cases)]
    frames.append(np.zeros(total, np.float32))                                              # silence: energy gate
    frames.append(np.random.default_rng(9).normal(0, 0.2, total).astype(np.float32))        # noise only: outlier / threshold tests
    x = np.stack(frames)
    llr, n, ds, cfo, ph = dem.receive_batch(x)
    found = 0
    for b in range(len(x)):
        ol, ods, ocfo, oph = O.dpsk_receive(mod, 384, x[b])
        assert int(ds[b]) == ods, (b, ds[b], ods)
        assert words(cfo[b])[()] == words(np.float32(ocfo))[()] and words(ph[b])[()] == words(np.float32(oph))[()], (b, cfo[b], ocfo, ph[b], oph)
        want = ol[:648]
        assert int(n[b]) == len(want), (b, n[b], len(want))
        assert (words(llr[b, :len(want)]) == words(want)).all(), b
        found += ods > 0
        if R.available() and b < 4:
            rl, rds, rcfo, rph = R.dpsk_receive(mod, 384, x[b])
            assert rds == ods and (words(rl[:648]) == words(want)).all()
    assert found >= 5
    d = dem.receive_batch(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    assert (d[2].cpu().numpy() == ds).all() and (words(d[0].cpu().numpy()) == words(llr)).all() and (d[1].cpu().numpy() == n).all()
    # short frames: below 1.5 preambles findPreamble gives up (:354-355)
    short = dem.receive_batch(x[:2, :39 * 384 + 100])
    assert (short[2] == -1).all() and (short[1] == 0).all()
    del ctx
