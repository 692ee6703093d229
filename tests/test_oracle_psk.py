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


def _words(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_golden_dpsk_acquisition(golden):
    """SURVEY §8f next-2 (DPSK half): the oracle's DPSKDemodulator::findPreamble + demodulateSoft restatement against
    vectors produced by the unmodified reference (tools/test_dpsk_snr.cpp:66-73 receive sequence)."""
    g = golden["dpsk_acquire"]
    for i in range(2):
        llr, ds, cfo, ph = O.dpsk_receive(2, 384, g[f"d{i}_rx"])
        assert ds == int(g[f"d{i}_info"][0])
        assert (_words(np.array([cfo, ph], np.float32)) == _words(g[f"d{i}_cfo_phase"])).all()
        want = g[f"d{i}_llr"]
        assert len(llr) == len(want) and (_words(llr) == _words(want)).all()


@pytest.mark.parametrize("mod", [0, 1, 2])
def test_dpsk_acquisition_matches_reference(mod):
    if not R.available():
        pytest.skip("needs oracle/_ref (the compiled reference)")
    rng = np.random.default_rng(40 + mod)
    found = 0
    for snr, lead in ((12.0, 0), (2.0, 411), (-5.0, 64), (-11.0, 700), (-18.0, 250), (-30.0, 0)):
        data = rng.integers(0, 256, 20, dtype=np.uint8)
        tx = R.dpsk_tx(mod, 384, R.ldpc_encode(R.R1_4, data), 0)
        tx = (tx * (np.float32(0.5) / np.abs(tx).max())).astype(np.float32)
        w = np.concatenate([np.zeros(lead, np.float32), tx])
        p = float(np.mean(tx.astype(np.float64) ** 2))
        rx = (w + rng.normal(0.0, np.sqrt(p / 10 ** (snr / 10)), len(w))).astype(np.float32)
        rl, rds, rcfo, rph = R.dpsk_receive(mod, 384, rx)
        ol, ods, ocfo, oph = O.dpsk_receive(mod, 384, rx)
        assert rds == ods and (_words(np.array([rcfo, rph], np.float32)) == _words(np.array([ocfo, oph], np.float32))).all(), (snr, lead)
        assert len(rl) == len(ol) and (_words(rl) == _words(ol)).all(), (snr, lead)
        found += rds > 0
    assert found >= 3


def test_golden_mcdpsk_got_chirp(golden):
    """SURVEY §8a row a16 (second half): the oracle's processGotChirp restatement -- Hilbert-FIR CFO correction, processTraining,
    5 Hz false-positive rule, setReference, demodulateSoft -- against vectors produced by the unmodified reference."""
    g = golden["mcdpsk_chirp"]
    for i in range(2):
        cfo, after = (float(v) for v in g[f"m{i}_cfo"])
        llr, oafter = O.mcdpsk_got_chirp(8, g[f"m{i}_rx"], cfo)
        want = g[f"m{i}_llr"]
        assert len(llr) == len(want) and (_words(llr) == _words(want)).all(), i
        assert (_words(np.float32(oafter)) == _words(np.float32(after))).all()
        assert bool(g[f"m{i}_ready"][0]) == (len(want) > 0)


@pytest.mark.parametrize("nc", [5, 8, 13, 20])
def test_mcdpsk_got_chirp_matches_reference(nc):
    if not R.available():
        pytest.skip("needs oracle/_ref (the compiled reference)")
    from projectultra_b200 import capi
    cfg = capi.mcdpsk_config(nc, 2)
    rng = np.random.default_rng(500 + nc)
    accepted = 0
    for snr, cfo in ((18.0, 0.0), (9.0, 0.1), (9.0, 0.12), (4.0, -1.7), (14.0, 9.0), (0.0, 0.3)):
        tx = capi.mcdpsk_tx(cfg, capi.ldpc_encode(capi.R1_4, rng.integers(0, 256, 20, dtype=np.uint8)))
        p = float(np.mean(tx.astype(np.float64) ** 2))
        rx = (tx + rng.normal(0.0, np.sqrt(p / 10 ** (snr / 10)), len(tx))).astype(np.float32)
        rl, rr, rc = R.mcdpsk_got_chirp(nc, rx, cfo)
        ol, oc = O.mcdpsk_got_chirp(nc, rx, cfo)
        assert len(rl) == len(ol) and (_words(rl) == _words(ol)).all() and (_words(np.float32(rc)) == _words(np.float32(oc))).all(), (snr, cfo)
        accepted += len(ol) > 0
    assert accepted >= 3


def test_golden_mcdpsk_chirp_receive(golden):
    """SURVEY §8f next-2 on MC-DPSK: the oracle's MCDPSKWaveform receive sequence (detectDualChirp -> training start two chirps and
    two gaps behind the up chirp -> setChirpDetected(cfo) -> process -> getSoftBits) against vectors from the unmodified reference."""
    g = golden["mcdpsk_chirp"]
    for i in range(2):
        llr, info, f, after = O.mcdpsk_chirp_receive(8, g[f"r{i}_rx"])
        want = g[f"r{i}_llr"]
        assert (info == g[f"r{i}_info"]).all(), i
        assert (_words(f) == _words(g[f"r{i}_f"])).all() and (_words(np.float32(after)) == _words(g[f"r{i}_after"][0])).all(), i
        assert len(llr) == len(want) and (_words(llr) == _words(want)).all(), i


@pytest.mark.parametrize("nc", [5, 8, 20])
def test_mcdpsk_chirp_receive_matches_reference(nc):
    if not R.available():
        pytest.skip("needs oracle/_ref (the compiled reference)")
    from projectultra_b200 import capi
    from mcframes import mcdpsk_chirp_frame
    cfg = capi.mcdpsk_config(nc, 2)
    rng = np.random.default_rng(700 + nc)
    got = 0
    for snr, lead, cfo, total in ((14.0, 500, 0.0, None), (6.0, 0, 7.3, None), (10.0, 1234, -22.0, None), (-14.0, 100, 0.0, None),
                                  (12.0, 300, 3.0, 60000), (12.0, 0, 0.0, 57600 + 9 * 512)):
        rx = mcdpsk_chirp_frame(cfg, rng, snr, lead, cfo, total)
        rl, ri, rf, ra = R.mcdpsk_chirp_receive(nc, rx)
        ol, oi, of, oa = O.mcdpsk_chirp_receive(nc, rx)
        assert (ri == oi).all() and (_words(rf) == _words(of)).all() and (_words(np.float32(ra)) == _words(np.float32(oa))).all(), (snr, lead, cfo)
        assert len(rl) == len(ol) and (_words(rl) == _words(ol)).all(), (snr, lead, cfo)
        got += len(ol) > 0
    assert got >= 2      # the 5 Hz false-positive rule rejects some of the CFO-free frames at 20 carriers
