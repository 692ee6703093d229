"""Pins the plain-C restatement of the single- and multi-carrier DPSK demodulators (oracle/pu_oracle_psk.c) bit for
bit against the unmodified reference (oracle/_ref, src/psk/dpsk.hpp, src/psk/multi_carrier_dpsk.hpp) and against the
committed golden vectors generated from it."""
import numpy as np
import pytest

import oracleapi as O
import refapi as R


def same_bits(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return a.shape == b.shape and bool((a == b).all())


def noisy(x, snr_db, rng):
    p = float(np.mean(x.astype(np.float64) ** 2))
    return (x + rng.standard_normal(len(x)).astype(np.float32) * np.float32(np.sqrt(p / 10 ** (snr_db / 10)))).astype(np.float32)


@pytest.mark.ref
@pytest.mark.parametrize("mod", [0, 1, 2])
def test_sc_dpsk_vs_reference(mod):
    rng = np.random.default_rng(10 + mod)
    for sps, snr, layout in ((384, 12.0, 0), (384, -3.0, 0), (192, 5.0, 1), (384, 30.0, 2)):
        data = rng.integers(0, 256, 27, dtype=np.uint8)
        tx = R.dpsk_tx(mod, sps, data, layout)
        tx = (tx * np.float32(0.5 / np.abs(tx).max())).astype(np.float32)
        rx = noisy(tx, snr, rng)
        start = {0: 39 * sps, 1: sps, 2: 0}[layout]
        for ref_mode, cfo, ph in ((0, 0.0, 0.0), (1, 0.0, 0.0), (1, 3.7, -0.4), (0, 0.2, 0.02), (1, -11.0, 2.9)):
            if ref_mode == 1 and start == 0:
                continue
            want = R.dpsk_demod_soft_ex(mod, sps, rx, start, ref_mode, cfo, ph)
            got = O.dpsk_demod_soft(mod, sps, rx, start, ref_mode, cfo, ph)
            assert len(want) == (len(rx) - start) // sps * (mod + 1)
            assert same_bits(got, want), (mod, sps, snr, ref_mode, cfo)
    # silence and ragged tails
    z = np.zeros(5 * 384 + 17, np.float32)
    assert same_bits(O.dpsk_demod_soft(mod, 384, z, 384, 1), R.dpsk_demod_soft_ex(mod, 384, z, 384, 1))


@pytest.mark.ref
@pytest.mark.parametrize("nc,bits", [(8, 2), (5, 2), (13, 2), (20, 2), (3, 1), (10, 1)])
def test_mc_dpsk_vs_reference(nc, bits):
    rng = np.random.default_rng(100 + nc)
    for snr in (20.0, 3.0, -6.0):
        data = rng.integers(0, 256, 81, dtype=np.uint8)
        tx = R.mcdpsk_tx(nc, data, bits=bits)
        rx = noisy(tx, snr, rng)
        want, wcfo = R.mcdpsk_demod_soft(nc, rx, bits=bits)
        got, gcfo = O.mcdpsk_demod_soft(nc, rx, bits=bits)
        assert len(want) >= 648
        assert same_bits(got, want), (nc, bits, snr)
        assert np.float32(gcfo) == np.float32(wcfo)
    z = np.zeros(12 * 512 + 100, np.float32)
    assert same_bits(O.mcdpsk_demod_soft(nc, z, bits=bits)[0], R.mcdpsk_demod_soft(nc, z, bits=bits)[0])


@pytest.mark.parametrize("mod", [0, 1, 2])
def test_sc_dpsk_golden(golden, mod):
    g = golden["psk"]
    rx, sps = g[f"sc{mod}_rx"], 192
    assert same_bits(O.dpsk_demod_soft(mod, sps, rx, 9 * sps, 1), g[f"sc{mod}_llr_ref1"])
    assert same_bits(O.dpsk_demod_soft(mod, sps, rx, 9 * sps, 0), g[f"sc{mod}_llr_ref0"])
    assert same_bits(O.dpsk_demod_soft(mod, sps, rx, 9 * sps, 1, 7.25, -0.6), g[f"sc{mod}_llr_comp"])


@pytest.mark.parametrize("nc,bits", [(8, 2), (3, 2), (5, 1)])
def test_mc_dpsk_golden(golden, nc, bits):
    g = golden["psk"]
    llr, cfo = O.mcdpsk_demod_soft(nc, g[f"mc{nc}_rx"], bits=bits)
    assert same_bits(llr, g[f"mc{nc}_llr"]) and np.float32(cfo) == g[f"mc{nc}_cfo"][0]


def test_psk_tx_golden(golden):
    from projectultra_b200 import build, capi
    build.build()
    g = golden["psk"]
    for mod in (0, 1, 2):
        tx = capi.dpsk_tx(capi.dpsk_config(mod, 192), g[f"sc{mod}_data"], 0)
        assert same_bits(tx[: 41 * 192], g[f"sc{mod}_tx_head"])
    for nc, bits in ((8, 2), (3, 2), (5, 1)):
        assert same_bits(capi.mcdpsk_tx(capi.mcdpsk_config(nc, bits), g[f"mc{nc}_data"]), g[f"mc{nc}_tx"])
