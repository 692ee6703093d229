"""The reference tools' CFO injector (tools/test_iwaveform.cpp:67-118: FFT-Hilbert analytic signal, float phase recurrence) as the sweep
driver applies it to the clean TX audio: pu_tools_apply_cfo (csrc/tools_cfo.cpp, host code) against the same loop around the COMPILED
reference FFT class (oracle/ref_build/ref_harness.cpp: ref_tools_apply_cfo), bit for bit, and against what a frequency shift must do."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import refapi as R
import oracleapi as O


def signal(n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 48000.0
    return (0.4 * np.sin(2 * np.pi * 1000.0 * t) + 0.2 * np.sin(2 * np.pi * 2300.0 * t + 1.0) + 0.05 * rng.standard_normal(n)).astype(np.float32)


@pytest.mark.skipif(not R.available(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("n,cfo", [(127, 30.0), (128, 30.0), (5000, 30.0), (65536, -50.0), (65537, 12.5), (9000, 0.0005), (70000, 50.0)])
def test_matches_the_loop_around_the_compiled_reference_fft(n, cfo):
    from projectultra_b200 import capi
    x = signal(n, n)
    got, want = capi.tools_apply_cfo(x, cfo), R.tools_apply_cfo(x, cfo)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    if n < 128 or abs(cfo) < 0.001:
        assert (got.view(np.uint32) == x.view(np.uint32)).all()           # untouched (:68)


@pytest.mark.parametrize("n,cfo", [(127, 30.0), (4096, 30.0), (5000, -50.0), (65537, 12.5), (9000, 0.0005)])
def test_matches_the_plain_c_oracle(n, cfo):
    """The same against oracle/pu_oracle_ofdm.c: orc_tools_apply_cfo (which the test above pins to the compiled reference where it is present)."""
    from projectultra_b200 import capi
    x = signal(n, 2 * n + 1)
    got, want = capi.tools_apply_cfo(x, cfo), O.tools_apply_cfo(x, cfo)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    if R.available():
        assert (R.tools_apply_cfo(x, cfo).view(np.uint32) == want.view(np.uint32)).all()


def test_shifts_every_component_by_the_offset():
    from projectultra_b200 import capi
    n, cfo = 48000, 30.0
    x = signal(n, 3)
    y = capi.tools_apply_cfo(x, cfo)
    sx, sy = np.abs(np.fft.rfft(x.astype(np.float64))), np.abs(np.fft.rfft(y.astype(np.float64)))
    assert abs(int(sx.argmax()) - 1000) <= 1 and abs(int(sy.argmax()) - 1030) <= 1        # 1 Hz bins
    assert sy[2330 - 2:2330 + 3].max() > 0.4 * sy.max() and sy[2300] < 0.05 * sy.max()
    assert abs(float(np.mean(y.astype(np.float64) ** 2)) / float(np.mean(x.astype(np.float64) ** 2)) - 1.0) < 0.02
