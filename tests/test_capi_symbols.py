"""CPU-only checks of the C-ABI shared library: it builds, loads, exports every symbol declared in
include/pu/*.h, fails loudly without a GPU, and its host-side helpers (encoder, interleaver tables) agree
with the oracle."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

import oracleapi as O
import refapi as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from projectultra_b200 import build, capi
    build.build()
    return capi


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "pu", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"PU_API\s+[\w\s\*]+?\b(pu_\w+)\s*\(", src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported(capi):
    L = capi.lib()
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_init_fails_loudly_without_gpu(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.PuError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


@pytest.mark.parametrize("rate", [0, 2, 3, 4, 5])
def test_host_encoder_matches_oracle(capi, rate):
    rng = np.random.default_rng(rate)
    for n in (1, R.RATE_K[rate] // 8, 100, 279):
        data = rng.integers(0, 256, n, dtype=np.uint8)
        assert (capi.ldpc_encode(rate, data) == O.ldpc_encode(rate, data)).all()


@pytest.mark.parametrize("rate", [0, 1, 2, 3, 4, 5, 6])
def test_host_frame_encoder_matches_oracle(capi, rate):
    """pu_frame_encode == v2::encodeFrameWithLDPC (the oracle's restatement is pinned to the reference in test_oracle_fec.py);
    R1/3 and R7/8 take the table's 27 / 20 bytes per codeword with the decoder's R1/2 fallback dimensions."""
    import v2frames as V
    rng = np.random.default_rng(60 + rate)
    for plen in (0, 1, 21, 100, 333):
        fr = V.data_frame(rng.integers(0, 256, plen, dtype=np.uint8), min(rate, 5))
        a, b = capi.frame_encode(rate, fr), O.frame_encode(rate, fr)
        assert a.shape == b.shape and (a == b).all(), (rate, plen)


def test_host_interleaver_tables_match_oracle(capi):
    x = np.arange(648, dtype=np.float32)
    for bps in (30, 60, 90, 118, 220, 708):
        perm, inv, step = capi.channel_interleaver_perm(bps)
        assert step == O.channel_interleaver_step(bps)
        y = np.zeros(648, np.float32)
        y[perm] = x
        assert (y == O.channel_interleave(bps, x)).all()
        z = np.zeros(648, np.float32)
        z[inv] = y
        assert (z == x).all()
    perm = capi.block_interleaver_perm(6, 108)
    y = np.zeros(648, np.float32)
    y[perm] = x
    assert (y == O.block_interleave(6, 108, x)).all()
