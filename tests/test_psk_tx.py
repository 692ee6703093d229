"""Host DPSK transmitters of the C ABI (pu_dpsk_tx, pu_mcdpsk_tx) against the unmodified reference's modulators:
bit-identical waveforms (same expressions, same host libm)."""
import numpy as np
import pytest

import refapi as R


@pytest.fixture(scope="module")
def capi():
    from projectultra_b200 import build, capi
    build.build()
    return capi


def same_bits(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return a.shape == b.shape and bool((a == b).all())


@pytest.mark.ref
@pytest.mark.parametrize("mod", [0, 1, 2])
def test_sc_dpsk_tx_matches_reference(capi, mod):
    rng = np.random.default_rng(mod)
    for sps in (384, 192, 1536):
        for layout in (0, 1, 2):
            data = rng.integers(0, 256, 81 if sps != 1536 else 9, dtype=np.uint8)
            got = capi.dpsk_tx(capi.dpsk_config(mod, sps), data, layout)
            assert same_bits(got, R.dpsk_tx(mod, sps, data, layout)), (mod, sps, layout)
    assert len(capi.dpsk_tx(capi.dpsk_config(mod, 384), np.zeros(81, np.uint8), 0)) == 39 * 384 + 384 * (648 // (mod + 1))


@pytest.mark.ref
@pytest.mark.parametrize("nc,bits", [(8, 2), (3, 2), (5, 1), (13, 2), (20, 2)])
def test_mc_dpsk_tx_matches_reference(capi, nc, bits):
    rng = np.random.default_rng(nc)
    data = rng.integers(0, 256, 81, dtype=np.uint8)
    got = capi.mcdpsk_tx(capi.mcdpsk_config(nc, bits), data)
    assert same_bits(got, R.mcdpsk_tx(nc, data, bits=bits)), (nc, bits)
