"""With the host libm routines restated on the device (csrc/ref_math.cuh) the CUDA receive path is expected to be
bit-identical to the oracle, not merely within tolerance.  This test measures that over every modulation and
requires >= 99.99 % identical LLR words and identical hard decisions; the remainder is bounded by the 1e-4 rule."""
import numpy as np
import pytest

import oracleapi as O
import refapi as R
from golden.make_golden import awgn
from test_ofdm_gpu import llr_mismatches, to_capi_cfg

pytestmark = pytest.mark.gpu
MODS = [R.DBPSK, R.DQPSK, R.D8PSK, R.BPSK, R.QPSK, R.QAM16, R.QAM32, R.QAM64, R.QAM256]


@pytest.mark.parametrize("preset", ["m1", "m3"])
def test_llr_words_identical(preset):
    from projectultra_b200 import capi
    ctx = capi.Context(0)
    total = same = 0
    for mod in MODS:
        rate = R.R1_2 if preset == "m1" else R.R3_4
        cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
        dem = capi.OfdmDemodulator(ctx, to_capi_cfg(cfg))
        frames, cfos, phs = [], [], []
        for i in range(24):
            rng = np.random.default_rng(mod * 1000 + i)
            data = rng.integers(0, 256, 40 if preset == "m1" else 60, dtype=np.uint8)
            tx = O.ofdm_tx(cfg, O.ldpc_encode(rate, data), 0)
            frames.append(awgn(tx, float(rng.uniform(-2, 32)), rng))
            use_cfo = i % 3 == 2
            cfos.append(float(rng.uniform(-40, 40)) if use_cfo else 0.0)
            phs.append(float(rng.uniform(-3, 3)) if use_cfo else 0.0)
        frames = np.stack(frames)
        n = dem.n_llr(frames.shape[1])
        ref, _ = O.ofdm_presynced_batch(cfg, frames, n, 2, 2, np.array(cfos, np.float32), np.array(phs, np.float32))
        got, _, _ = dem.presynced_batch(frames, 2, np.array(cfos, np.float32), np.array(phs, np.float32))
        eq = got.view(np.uint32) == ref.view(np.uint32)
        total += eq.size
        same += int(eq.sum())
        assert len(llr_mismatches(got.ravel(), ref.ravel())) == 0, mod
        assert (np.signbit(got) == np.signbit(ref)).all(), mod
    frac = same / total
    print(f"\n[{preset}] bit-identical LLR words: {same}/{total} = {frac:.6f}")
    assert frac >= 0.9999
    del ctx
