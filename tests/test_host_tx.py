"""CPU-only: the product's host transmitter (pu_ofdm_tx) and channel sigma helper are bit-identical to the oracle."""
import numpy as np
import pytest

import oracleapi as O
import refapi as R


@pytest.fixture(scope="module")
def ls():
    from projectultra_b200 import build, linksim
    build.build()
    return linksim


def to_capi_cfg(cfg):
    from projectultra_b200 import capi
    return capi.ModemConfig.from_buffer_copy(bytes(cfg))


@pytest.mark.parametrize("preset", ["m1", "m3"])
@pytest.mark.parametrize("mod", [R.DBPSK, R.DQPSK, R.D8PSK, R.BPSK, R.QPSK, R.QAM16, R.QAM32, R.QAM64, R.QAM256])
def test_tx_matches_oracle(ls, preset, mod):
    rate = R.R1_2 if preset == "m1" else R.R3_4
    cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
    rng = np.random.default_rng(mod)
    for nbytes in (1, 40, 81, 200):
        data = rng.integers(0, 256, nbytes, dtype=np.uint8)
        for layout in (0, 1):
            a = ls.ofdm_tx(to_capi_cfg(cfg), data, layout)
            b = O.ofdm_tx(cfg, data, layout)
            assert a.shape == b.shape and (a.view(np.uint32) == b.view(np.uint32)).all(), (nbytes, layout)


def test_tx_cfo_and_golden(ls, golden):
    g = golden["ofdm"]
    cfg = R.config_m1(R.DQPSK, R.R1_2)
    assert (ls.ofdm_tx(to_capi_cfg(cfg), g["tx_m1_dqpsk_cw"], 0).view(np.uint32) == g["tx_m1_dqpsk_l0"].view(np.uint32)).all()
    assert (ls.ofdm_tx(to_capi_cfg(cfg), g["tx_m1_dqpsk_cw"], 1).view(np.uint32) == g["tx_m1_dqpsk_l1"].view(np.uint32)).all()
    cfg.tx_cfo_hz = 12.5
    data = np.arange(81, dtype=np.uint8)
    assert (ls.ofdm_tx(to_capi_cfg(cfg), data, 0).view(np.uint32) == O.ofdm_tx(cfg, data, 0).view(np.uint32)).all()


def test_noise_std_conventions(ls):
    import ctypes as C
    L = O.lib()
    L.orc_channel_noise_std.restype = C.c_float
    rng = np.random.default_rng(0)
    x = rng.standard_normal(7332).astype(np.float32) * 0.3
    for conv in (0, 1):
        for snr in (-11.0, 0.0, 17.5, 40.0):
            want = L.orc_channel_noise_std(x.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(len(x)), C.c_float(snr), conv)
            assert ls.channel_noise_std(x, snr, conv) == want
