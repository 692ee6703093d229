"""the batched transmitter on the GPU (LDPC encode + OFDM modulate) against the host transmitter, which is
itself pinned bit-for-bit to the reference's LDPCEncoder / OFDMModulator (tests/test_host_tx.py, tests/test_oracle_ofdm.py)
and here again against the plain-C oracle.  Every waveform sample must be bit-identical."""
import numpy as np
import pytest

import refapi as R
import oracleapi as O

pytestmark = pytest.mark.gpu

MODS = [R.DBPSK, R.DQPSK, R.D8PSK, R.BPSK, R.QPSK, R.QAM16, R.QAM32, R.QAM64, R.QAM256]


def words(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("preset", ["m1", "m3"])
@pytest.mark.parametrize("mod", MODS)
def test_tx_batch_bit_identical(preset, mod):
    import torch
    from projectultra_b200 import capi
    ctx = capi.Context(0)
    rng = np.random.default_rng(100 + mod)
    for rate, nbytes in ((R.R1_2, 40), (R.R3_4, 60), (R.R1_4, 20), (R.R5_6, 67)):
        cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
        dem = capi.OfdmDemodulator(ctx, capi.ModemConfig.from_buffer_copy(bytes(cfg)))
        enc = capi.LdpcDecoder(ctx, rate)
        payload = rng.integers(0, 256, (5, nbytes), dtype=np.uint8)
        payload[3] = 0
        payload[4] = 255
        for layout in (0, 1):
            want = np.stack([O.ofdm_tx(cfg, O.ldpc_encode(rate, p), layout) for p in payload])
            got = dem.tx_batch(enc, payload, layout=layout)
            assert got.shape == want.shape, (got.shape, want.shape)
            assert (words(got) == words(want)).all(), (preset, mod, rate, layout, np.flatnonzero(words(got) != words(want))[:8])
            dgot = dem.tx_batch(enc, torch.from_numpy(payload).cuda(), layout=layout, peak=0.5)
            torch.cuda.synchronize()
            scaled = np.stack([(w * (np.float32(0.5) / np.abs(w).max())).astype(np.float32) for w in want])
            assert (words(dgot.cpu().numpy()) == words(scaled)).all(), (preset, mod, rate, layout, "peak")
        if R.available():
            ref = R.ofdm_tx(cfg, R.ldpc_encode(rate, payload[0]), 1)
            assert (words(dem.tx_batch(enc, payload[:1], layout=1)[0]) == words(ref)).all()
    del ctx
