"""Pins the C oracle's OFDM layer (oracle/pu_oracle_ofdm.c) against golden vectors from the unmodified
reference and, when oracle/_ref is built, bit-for-bit against the reference stage by stage.  CPU only."""
import numpy as np
import pytest

import oracleapi as O
import refapi as R
from golden.make_golden import OFDM_CASES, awgn


def cfg_from(g, name):
    return R.ModemConfig.from_buffer_copy(bytes(g[name + "_cfg"]))


def same_bits(a, b):
    a = np.ascontiguousarray(a).view(np.uint32)
    b = np.ascontiguousarray(b).view(np.uint32)
    return a.shape == b.shape and bool((a == b).all())


@pytest.mark.parametrize("case", [c[0] for c in OFDM_CASES])
def test_golden_ofdm_rx(golden, case):
    g = golden["ofdm"]
    cfg = cfg_from(g, case)
    cm, cfo, ph = g[case + "_cfo"]
    st = O.ofdm_presynced_stages(cfg, g[case + "_rx"], 2, int(cm), float(cfo), float(ph))
    assert same_bits(st["llr"], g[case + "_llr"])
    assert same_bits(st["scalars"], g[case + "_scalars"])
    assert same_bits(st["h"][-1], g[case + "_h_last"])
    info, ok, it = O.ldpc_decode_soft(cfg.code_rate, st["llr"][:648])
    assert (info == g[case + "_info"]).all() and [int(ok), it] == list(g[case + "_ok"])


def test_golden_ofdm_tx(golden):
    g = golden["ofdm"]
    cfg = R.config_m1(R.DQPSK, R.R1_2)
    assert same_bits(O.ofdm_tx(cfg, g["tx_m1_dqpsk_cw"], 0), g["tx_m1_dqpsk_l0"])
    assert same_bits(O.ofdm_tx(cfg, g["tx_m1_dqpsk_cw"], 1), g["tx_m1_dqpsk_l1"])
    assert len(g["tx_m1_dqpsk_l0"]) == 7332 and len(g["tx_m1_dqpsk_l1"]) == 10124


def test_golden_dsp(golden):
    g = golden["misc"]
    assert same_bits(O.nco(1500, 48000, 2048), g["nco_1500_48000"])
    assert same_bits(O.fft(g["fft512_in"]), g["fft512_out"])
    x = g["fft512_in"]
    assert np.abs(O.fft(O.fft(x), inverse=True) - x).max() < 1e-5       # tests/test_fft.cpp round trip
    tone = np.exp(2j * np.pi * 8 * np.arange(512) / 512).astype(np.complex64)
    assert int(np.argmax(np.abs(O.fft(tone)))) == 8                      # tests/test_fft.cpp tone at bin 8


def test_demapper_sign_conventions():
    # tests/test_comprehensive_modem.cpp:270-377: + LLR <=> bit 0
    assert O.soft_demap(R.BPSK, -1 + 0j)[0] > 0 and O.soft_demap(R.BPSK, 1 + 0j)[0] < 0
    q = 0.70710678
    assert (O.soft_demap(R.QPSK, complex(-q, -q)) > 0).all() and (O.soft_demap(R.QPSK, complex(q, q)) < 0).all()
    for bits, ang in ((0b00, 0), (0b01, 90), (0b10, 180), (0b11, 270)):  # modulator.cpp:414-421
        l = O.soft_demap(R.DQPSK, np.exp(1j * np.deg2rad(ang)), 1 + 0j)
        assert [int(v < 0) for v in l] == [bits >> 1, bits & 1]
    assert (O.soft_demap(R.DQPSK, 1e-4 + 0j, 1e-3 + 0j) == 0).all()       # weak-signal gate, soft_demap.hpp:199-201
    assert O.soft_demap(R.QAM16, 0.01 + 0.0j, nv=10.0)[0] == -0.5          # clipLLR min magnitude, :22-29


@pytest.mark.ref
@pytest.mark.parametrize("mod", [R.DBPSK, R.DQPSK, R.D8PSK, R.BPSK, R.QPSK, R.QAM16, R.QAM32, R.QAM64, R.QAM256])
@pytest.mark.parametrize("preset", ["m1", "m3"])
def test_vs_reference_stages(mod, preset):
    rate = R.R1_2 if preset == "m1" else R.R3_4
    cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
    for i, snr in enumerate((4.0, 16.0, 30.0)):
        rng = np.random.default_rng(mod * 100 + i)
        data = rng.integers(0, 256, 40 if preset == "m1" else 60, dtype=np.uint8)
        cw = R.ldpc_encode(rate, data)
        tx = R.ofdm_tx(cfg, cw, 0)
        assert same_bits(tx, O.ofdm_tx(cfg, cw, 0))
        rx = awgn(tx, snr, rng) if i != 1 else R.watterson(tx, snr, 1.0, 0.5, seed=42 + mod)
        cm, cfo, ph = (1, 0.0, 0.0) if i != 2 else (2, 7.25, 1.1)
        a = R.ofdm_presynced_stages(cfg, rx, 2, cm, cfo, ph)
        b = O.ofdm_presynced_stages(cfg, rx, 2, cm, cfo, ph)
        for key in ("carriers", "lts_bins", "h_lts", "bins", "h", "eq", "nv", "scalars", "llr"):
            assert same_bits(a[key], b[key]), (key, snr)


@pytest.mark.ref
def test_vs_reference_sc_preamble_tx():
    for mod in (R.DQPSK, R.QAM16):
        cfg = R.config_m1(mod, R.R1_2)
        cw = R.ldpc_encode(R.R1_2, np.arange(40, dtype=np.uint8))
        assert same_bits(R.ofdm_tx(cfg, cw, 1), O.ofdm_tx(cfg, cw, 1))


def _same_words(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return a.shape == b.shape and bool((a == b).all())


def test_golden_acquisition(golden):
    """the oracle's restatement of OFDMDemodulator::process (Schmidl-Cox search, coarse CFO, LTS timing,
    SYNCED symbol loop) against vectors produced by the unmodified reference fed in 960-sample chunks and in one piece."""
    from golden.make_golden import ACQ_CASES
    g = golden["acquire"]
    for i in range(len(ACQ_CASES)):
        cfg = R.ModemConfig.from_buffer_copy(bytes(g[f"a{i}_cfg"]))
        rx = g[f"a{i}_rx"]
        for ci, chunk in enumerate((960, len(rx))):
            llr, synced, off, cfo, data_start, calls = O.ofdm_process(cfg, rx, chunk)
            want_sync, want_off = (int(v) for v in g[f"a{i}_c{ci}_info"])
            assert int(synced) == want_sync, (i, chunk)
            if not want_sync:
                assert len(llr) == 0
                continue
            assert off == want_off, (i, chunk, off, want_off)
            assert _same_words(np.float32(cfo), g[f"a{i}_c{ci}_cfo"][0]), (i, chunk, cfo)
            assert _same_words(llr, g[f"a{i}_c{ci}_llr"]), (i, chunk)


@pytest.mark.parametrize("preset,mod", [("m1", R.DQPSK), ("m1", R.D8PSK), ("m3", R.DQPSK)])
def test_process_path_matches_reference(preset, mod):
    """Same restatement against the compiled reference itself on fresh frames across the synchronisation threshold."""
    if not R.available():
        pytest.skip("needs oracle/_ref (the compiled reference)")
    rate = R.R1_2 if preset == "m1" else R.R3_4
    cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
    rng = np.random.default_rng(31 + mod)
    n_sync = 0
    for snr in (30.0, 24.0, 20.0, 17.0, 15.0, 11.0):
        data = rng.integers(0, 256, 40 if preset == "m1" else 60, dtype=np.uint8)
        rx = awgn(O.ofdm_tx(cfg, O.ldpc_encode(rate, data), 1), snr, rng)
        for chunk in (960, 2500):
            rl, rs, roff, rcfo = R.ofdm_process_info(cfg, rx, chunk)
            ol, osync, ooff, ocfo, _, _ = O.ofdm_process(cfg, rx, chunk)
            assert rs == osync, (snr, chunk)
            if rs:
                n_sync += 1
                assert roff == ooff and _same_words(np.float32(rcfo), np.float32(ocfo)) and _same_words(rl, ol), (snr, chunk)
    assert n_sync >= 4


def test_golden_chirp_sync(golden):
    """SURVEY §8f next-2 (chirp half): the oracle's sync::ChirpSync::detectDualChirp + OFDMChirpWaveform receive glue against
    vectors produced by the unmodified reference (tools/test_iwaveform.cpp:127-160 sequence), with and without TX CFO."""
    g = golden["chirp"]
    base = R.config_m1(R.DQPSK, R.R1_2)
    for i in range(2):
        llr, info, cfo = O.ofdm_chirp_receive(base, g[f"c{i}_rx"])
        assert (info.astype(np.int64) == g[f"c{i}_info"]).all(), (i, info, g[f"c{i}_info"])
        assert _same_words(np.float32(cfo), g[f"c{i}_cfo"][0])
        assert _same_words(llr, g[f"c{i}_llr"]), i


def test_chirp_sync_matches_reference():
    if not R.available():
        pytest.skip("needs oracle/_ref (the compiled reference)")
    assert _same_words(O.chirp_generate(48000.0, 7.5), R.chirp_generate(48000.0, 7.5))
    rng = np.random.default_rng(77)
    base = R.config_m1(R.DQPSK, R.R1_2)
    found = 0
    for snr, lead, tx_cfo in ((20.0, 0, 0.0), (8.0, 2100, -20.0), (0.0, 640, 5.0), (-9.0, 3000, 0.0), (-25.0, 10, 0.0)):
        cfg = R.config_m1(R.DQPSK, R.R1_2)
        cfg.tx_cfo_hz = tx_cfo
        body = O.ofdm_tx(cfg, O.ldpc_encode(R.R1_2, rng.integers(0, 256, 40, dtype=np.uint8)), 0)
        w = np.concatenate([np.zeros(lead, np.float32), O.chirp_generate(48000.0, tx_cfo), body, np.zeros(900, np.float32)])
        p = float(np.mean(body.astype(np.float64) ** 2))
        rx = (w + rng.normal(0.0, np.sqrt(p / 10 ** (snr / 10)), len(w))).astype(np.float32)
        rl, ri, rc = R.ofdm_chirp_receive(base, rx)
        ol, oi, oc = O.ofdm_chirp_receive(base, rx)
        assert (ri == oi).all() and _same_words(np.float32(rc), np.float32(oc)) and _same_words(rl, ol), (snr, lead, tx_cfo, ri, oi)
        found += int(oi[0])
    assert found >= 3


def test_training_cfo_path_against_golden(golden):
    """reset(); processPresynced(span, 2) without setFrequencyOffset -> estimateCFOFromTraining (ofdm_sync.cpp:278-380): the oracle's
    restatement reproduces the reference's soft bits and adopted CFO bit for bit (vectors generated from the compiled reference)."""
    g = golden["training_cfo"]
    n = len([k for k in g.files if k.endswith("_rx")])
    assert n >= 6
    for i in range(n):
        cfg = R.ModemConfig.from_buffer_copy(bytes(g["c%d_cfg" % i]))
        rx = g["c%d_rx" % i]
        llr, snr, fc = O.ofdm_presynced(cfg, rx, 2, 0)
        want = g["c%d_llr" % i]
        assert len(llr) == len(want) and (llr.view(np.uint32) == want.view(np.uint32)).all(), i
        assert np.float32(fc).view(np.uint32) == g["c%d_final_cfo" % i].view(np.uint32), i


@pytest.mark.ref
def test_training_cfo_path_against_compiled_reference():
    from golden.make_golden import awgn
    for k, (mod, preset, tx_cfo, snr) in enumerate(((R.DQPSK, "m1", 0.0, 25.0), (R.DQPSK, "m1", 12.0, 0.0), (R.QAM16, "m1", 3.0, 8.0),
                                                    (R.QAM32, "m3", -7.5, 12.0), (R.DBPSK, "m3", 3.0, 4.0), (R.QPSK, "m1", -7.5, 10.0))):
        rate = R.R1_2 if preset == "m1" else R.R3_4
        cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
        cfg.tx_cfo_hz = tx_cfo
        rng = np.random.default_rng(900 + k)
        rx = awgn(O.ofdm_tx(cfg, O.ldpc_encode(rate, rng.integers(0, 256, 40, dtype=np.uint8)), 0), snr, rng)
        a, _, fa = R.ofdm_presynced(cfg, rx, 2, 0)
        b, _, fb = O.ofdm_presynced(cfg, rx, 2, 0)
        assert len(a) == len(b) and (a.view(np.uint32) == b.view(np.uint32)).all(), k
        assert np.float32(fa).view(np.uint32) == np.float32(fb).view(np.uint32), k
        # one training symbol: no estimate, CFO 0 (ofdm_sync.cpp:279-282)
        a1, _, f1 = R.ofdm_presynced(cfg, rx, 1, 0)
        b1, _, g1 = O.ofdm_presynced(cfg, rx, 1, 0)
        assert (a1.view(np.uint32) == b1.view(np.uint32)).all() and f1 == g1


def test_float_only_two_pi_wrap_equals_the_double_wrap():
    """The CFO rotator's phase wrap (channel_equalizer.cpp:47-50: `phase -= 2 * M_PI` on a float, i.e. float((double)ph - 2 pi)) as the
    warp kernel evaluates it without FP64 (csrc/ofdm_demod.cu: rot_step): fl(fl(ph -+ two_hi) -+ two_lo).  Every float of [pi, 4.2),
    both signs -- beyond that the kernel takes the double path."""
    pi_hi = np.float32(3.14159274101257324)
    bits = np.arange(pi_hi.view(np.uint32), np.float32(4.2).view(np.uint32), dtype=np.uint32)
    ph = bits.view(np.float32)
    assert len(ph) == 4019851
    two_hi, two_lo = np.float32(6.2831854820251465), np.float32(-1.7484555e-7)
    ref = (ph.astype(np.float64) - 2.0 * np.pi).astype(np.float32)
    got = (ph - two_hi) - two_lo
    assert (ref.view(np.uint32) == got.view(np.uint32)).all()
    ref = ((-ph).astype(np.float64) + 2.0 * np.pi).astype(np.float32)
    got = ((-ph) + two_hi) + two_lo
    assert (ref.view(np.uint32) == got.view(np.uint32)).all()
