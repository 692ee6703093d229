#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libpu_ref.so, built by
oracle/ref_build/Makefile from /root/reference).  Run in the build container (where /root/reference
exists); the vectors then travel with the repo so that the oracle and the CUDA path can be pinned on
boxes that have no reference checkout.

    python tests/golden/make_golden.py
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refapi as R  # noqa: E402


def ldpc_vectors():
    rng = np.random.default_rng(20261017)
    out = {}
    # Eb/N0-ish operating points: easy / waterfall / stress per rate (sigma of BPSK AWGN)
    sig = {R.R1_4: (0.7, 1.15, 1.4), R.R1_2: (0.5, 0.72, 0.9), R.R2_3: (0.45, 0.6, 0.8),
           R.R3_4: (0.4, 0.58, 0.8), R.R5_6: (0.4, 0.6, 0.8)}
    for rate, k in R.RATE_K.items():
        data = rng.integers(0, 256, (k + 7) // 8, dtype=np.uint8)
        cw = R.ldpc_encode(rate, data[:k // 8])
        bits = np.unpackbits(cw)[:648].astype(np.float32)
        llrs = []
        for s in sig[rate]:
            y = (1 - 2 * bits)[None, :] + s * rng.standard_normal((6, 648)).astype(np.float32)
            l = np.clip(2 * y / s ** 2, -10, 10).astype(np.float32)
            l = np.where(np.abs(l) < 0.5, np.where(l >= 0, 0.5, -0.5), l).astype(np.float32)
            llrs.append(l)
        # demod-realistic: exactly +-10 with iid sign flips -> all-tie minima
        for p in (0.02, 0.06, 0.13):
            flip = rng.random((4, 648)) < p
            l = np.where((bits[None, :] > 0) ^ flip, -10.0, 10.0).astype(np.float32)
            llrs.append(l)
        # erasures and unclamped large values
        l = llrs[0][:2].copy()
        l[:, ::7] = 0.0
        l[:, 5::11] *= 9.0
        llrs.append(l)
        llr = np.concatenate(llrs, 0)
        info, ok, it = R.ldpc_decode_batch(rate, llr)
        out[f"r{rate}_data"] = data
        out[f"r{rate}_cw"] = cw
        out[f"r{rate}_llr"] = llr
        out[f"r{rate}_info"] = info
        out[f"r{rate}_ok"] = ok
        out[f"r{rate}_iters"] = it
        # multi-block + partial block decodeSoft (ldpc_decoder.cpp:283-428)
        mb = np.concatenate([llr[0], llr[7], llr[1][:300]])
        o, okm, itm = R.ldpc_decode_soft(rate, mb)
        out[f"r{rate}_mb_llr"] = mb
        out[f"r{rate}_mb_out"] = o
        out[f"r{rate}_mb_ok"] = np.array([okm, itm], np.int32)
    np.savez_compressed(os.path.join(HERE, "ldpc_golden.npz"), **out)


def awgn(x, snr_db, rng):
    p = float((x.astype(np.float64) ** 2).mean())
    return (x + np.sqrt(p / 10 ** (snr_db / 10)) * rng.standard_normal(len(x))).astype(np.float32)


OFDM_CASES = [
    # name, preset, mod, rate, payload, snr_db, (cfo_mode,cfo,phase), channel(delay_ms,doppler)|None
    ("m1_dqpsk_awgn25", "m1", R.DQPSK, R.R1_2, 40, 25.0, (1, 0.0, 0.0), None),
    ("m1_dqpsk_awgn0", "m1", R.DQPSK, R.R1_2, 40, 0.0, (1, 0.0, 0.0), None),
    ("m1_dbpsk_awgn3", "m1", R.DBPSK, R.R1_4, 20, 3.0, (1, 0.0, 0.0), None),
    ("m1_d8psk_awgn12", "m1", R.D8PSK, R.R1_2, 40, 12.0, (1, 0.0, 0.0), None),
    ("m1_dqpsk_cfo", "m1", R.DQPSK, R.R1_2, 40, 18.0, (2, 12.5, 0.7), None),
    ("m1_dqpsk_flutter", "m1", R.DQPSK, R.R1_2, 40, 15.0, (1, 0.0, 0.0), (0.5, 10.0)),
    ("m1_bpsk_awgn6", "m1", R.BPSK, R.R1_2, 40, 6.0, (1, 0.0, 0.0), None),
    ("m1_qpsk_poor", "m1", R.QPSK, R.R1_2, 40, 15.0, (1, 0.0, 0.0), (2.0, 1.0)),
    ("m1_qam16_awgn14", "m1", R.QAM16, R.R1_2, 40, 14.0, (1, 0.0, 0.0), None),
    ("m1_qam16_cfo", "m1", R.QAM16, R.R1_2, 40, 20.0, (2, -31.0, -2.9), None),
    ("m1_qam64_awgn22", "m1", R.QAM64, R.R3_4, 60, 22.0, (1, 0.0, 0.0), None),
    ("m3_dqpsk_awgn10", "m3", R.DQPSK, R.R3_4, 60, 10.0, (1, 0.0, 0.0), None),
    ("m3_qam16_awgn12", "m3", R.QAM16, R.R3_4, 60, 12.0, (1, 0.0, 0.0), None),
    ("m3_qam32_awgn14", "m3", R.QAM32, R.R3_4, 60, 14.0, (1, 0.0, 0.0), None),
    ("m3_qam32_good25", "m3", R.QAM32, R.R3_4, 60, 25.0, (1, 0.0, 0.0), (0.5, 0.1)),
]


def ofdm_vectors():
    out = {}
    for i, (name, preset, mod, rate, nbytes, snr, (cm, cfo, ph), chan) in enumerate(OFDM_CASES):
        rng = np.random.default_rng(1000 + i)
        cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
        data = rng.integers(0, 256, nbytes, dtype=np.uint8)
        cw = R.ldpc_encode(rate, data)
        tx = R.ofdm_tx(cfg, cw, 0)
        if chan is None:
            rx = awgn(tx, snr, rng)
        else:
            rx = R.watterson(tx, snr, chan[0], chan[1], seed=42 + i)
        st = R.ofdm_presynced_stages(cfg, rx, 2, cm, cfo, ph)
        info, ok, it = R.ldpc_decode_soft(rate, st["llr"][:648])
        out[name + "_cfg"] = np.frombuffer(bytes(cfg), dtype=np.uint8).copy()
        out[name + "_cfo"] = np.array([cm, cfo, ph], np.float32)
        out[name + "_data"] = data
        out[name + "_rx"] = rx.astype(np.float16).astype(np.float32) if False else rx
        out[name + "_llr"] = st["llr"]
        out[name + "_scalars"] = st["scalars"]
        out[name + "_h_last"] = st["h"][-1]
        out[name + "_info"] = info
        out[name + "_ok"] = np.array([ok, it], np.int32)
    # transmitter golden: first presynced frame and first S-C frame, M1 DQPSK R1/2
    rng = np.random.default_rng(7)
    cfg = R.config_m1(R.DQPSK, R.R1_2)
    data = rng.integers(0, 256, 40, dtype=np.uint8)
    cw = R.ldpc_encode(R.R1_2, data)
    out["tx_m1_dqpsk_cw"] = cw
    out["tx_m1_dqpsk_l0"] = R.ofdm_tx(cfg, cw, 0)
    out["tx_m1_dqpsk_l1"] = R.ofdm_tx(cfg, cw, 1)
    np.savez_compressed(os.path.join(HERE, "ofdm_golden.npz"), **out)


def misc_vectors():
    out = {}
    for bps in (30, 60, 90, 118, 220, 708):
        x = np.arange(648, dtype=np.float32)
        out[f"ci_{bps}"] = R.channel_interleave(bps, x)
    out["bi_6x108"] = R.block_interleave(6, 108, np.arange(648, dtype=np.float32))
    out["nco_1500_48000"] = R.nco(1500, 48000, 2048)
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(512) + 1j * rng.standard_normal(512)).astype(np.complex64)
    out["fft512_in"] = x
    out["fft512_out"] = R.fft(x)
    np.savez_compressed(os.path.join(HERE, "misc_golden.npz"), **out)


def psk_vectors():
    """Short single- and multi-carrier DPSK frames (kept small: noise does not compress)."""
    out = {}
    for mod in (0, 1, 2):
        rng = np.random.default_rng(5000 + mod)
        sps = 192
        data = rng.integers(0, 256, 6, dtype=np.uint8)
        tx = R.dpsk_tx(mod, sps, data, 0)
        tx = (tx * np.float32(0.5 / np.abs(tx).max())).astype(np.float32)
        rx = awgn(tx, 4.0 + 3 * mod, rng)
        rx = rx[30 * sps:]                      # keep 9 preamble symbols + data
        start = 9 * sps
        out[f"sc{mod}_data"] = data
        out[f"sc{mod}_rx"] = rx
        out[f"sc{mod}_llr_ref1"] = R.dpsk_demod_soft_ex(mod, sps, rx, start, 1)
        out[f"sc{mod}_llr_ref0"] = R.dpsk_demod_soft_ex(mod, sps, rx, start, 0)
        out[f"sc{mod}_llr_comp"] = R.dpsk_demod_soft_ex(mod, sps, rx, start, 1, 7.25, -0.6)
        out[f"sc{mod}_tx_head"] = R.dpsk_tx(mod, sps, data, 0)[: 41 * sps]
    for nc, bits in ((8, 2), (3, 2), (5, 1)):
        rng = np.random.default_rng(6000 + nc)
        data = rng.integers(0, 256, 12, dtype=np.uint8)
        tx = R.mcdpsk_tx(nc, data, bits=bits)
        rx = awgn(tx, 5.0, rng)
        llr, cfo = R.mcdpsk_demod_soft(nc, rx, bits=bits)
        out[f"mc{nc}_data"] = data
        out[f"mc{nc}_rx"] = rx
        out[f"mc{nc}_llr"] = llr
        out[f"mc{nc}_cfo"] = np.array([cfo], np.float32)
        out[f"mc{nc}_tx"] = tx
    np.savez_compressed(os.path.join(HERE, "psk_golden.npz"), **out)


ACQ_CASES = [("m1", R.DQPSK, R.R1_2, 40, 25.0), ("m1", R.DQPSK, R.R1_2, 40, 19.0), ("m1", R.D8PSK, R.R1_2, 40, 28.0),
             ("m1", R.DQPSK, R.R1_2, 40, 9.0), ("m3", R.DQPSK, R.R3_4, 60, 24.0)]


def acquire_vectors():
    """Schmidl-Cox path: OFDMDemodulator::process fed in 960-sample chunks (tools/test_mode_snr.cpp:65-70)
    and in one piece, then getSoftBits(); frames = generatePreamble() + modulate() over AWGN on mean frame power."""
    rng = np.random.default_rng(20261018)
    out = {}
    for i, (preset, mod, rate, nbytes, snr) in enumerate(ACQ_CASES):
        cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
        data = rng.integers(0, 256, nbytes, dtype=np.uint8)
        tx = R.ofdm_tx(cfg, R.ldpc_encode(rate, data), 1)
        rx = awgn(tx, snr, rng)
        out[f"a{i}_cfg"] = np.frombuffer(bytes(cfg), np.uint8)
        out[f"a{i}_rx"] = rx
        for chunk in (960, len(rx)):
            llr, synced, off, cfo = R.ofdm_process_info(cfg, rx, chunk)
            tag = f"a{i}_c{0 if chunk == 960 else 1}"
            out[tag + "_llr"] = llr
            out[tag + "_info"] = np.array([int(synced), off], np.int64)
            out[tag + "_cfo"] = np.array([cfo], np.float32)
    np.savez_compressed(os.path.join(HERE, "acquire_golden.npz"), **out)


def dpsk_acquire_vectors():
    """Barker acquisition (DPSK half): the receive sequence of tools/test_dpsk_snr.cpp:66-73 on two D8PSK
    frames (98k samples each), one found at 0 dB behind 300 samples of noise, one lost at -25 dB."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from projectultra_b200 import capi
    rng = np.random.default_rng(20261019)
    out = {}
    cfg = capi.dpsk_config(2, 384)
    for i, (snr, lead) in enumerate(((0.0, 300), (-25.0, 0))):
        data = rng.integers(0, 256, 20, dtype=np.uint8)
        tx = R.dpsk_tx(2, 384, R.ldpc_encode(R.R1_4, data), 0)
        tx = (tx * (np.float32(0.5) / np.abs(tx).max())).astype(np.float32)
        w = np.concatenate([np.zeros(lead, np.float32), tx])
        p = float(np.mean(tx.astype(np.float64) ** 2))
        rx = (w + rng.normal(0.0, np.sqrt(p / 10 ** (snr / 10)), len(w))).astype(np.float32)
        llr, ds, cfo, ph = R.dpsk_receive(2, 384, rx)
        out[f"d{i}_rx"] = rx
        out[f"d{i}_llr"] = llr
        out[f"d{i}_info"] = np.array([ds], np.int64)
        out[f"d{i}_cfo_phase"] = np.array([cfo, ph], np.float32)
    np.savez_compressed(os.path.join(HERE, "dpsk_acquire_golden.npz"), **out)


def chirp_vectors():
    """Dual-chirp sync (chirp half): tools/test_iwaveform.cpp:127-160 receive sequence on two OFDM_CHIRP frames
    (M1 DQPSK R1/2): +12.5 Hz TX CFO at 15 dB behind 1 200 samples of noise, and 0 Hz at -6 dB."""
    rng = np.random.default_rng(20261020)
    out = {}
    for i, (snr, lead, tx_cfo) in enumerate(((15.0, 1200, 12.5), (-6.0, 300, 0.0))):
        cfg = R.config_m1(R.DQPSK, R.R1_2)
        cfg.tx_cfo_hz = tx_cfo
        data = rng.integers(0, 256, 40, dtype=np.uint8)
        body = R.ofdm_tx(cfg, R.ldpc_encode(R.R1_2, data), 0)
        w = np.concatenate([np.zeros(lead, np.float32), R.chirp_generate(48000.0, tx_cfo), body, np.zeros(600, np.float32)])
        p = float(np.mean(body.astype(np.float64) ** 2))
        rx = (w + rng.normal(0.0, np.sqrt(p / 10 ** (snr / 10)), len(w))).astype(np.float32)
        base = R.config_m1(R.DQPSK, R.R1_2)
        llr, info, cfo = R.ofdm_chirp_receive(base, rx)
        out[f"c{i}_rx"] = rx
        out[f"c{i}_llr"] = llr
        out[f"c{i}_info"] = info.astype(np.int64)
        out[f"c{i}_cfo"] = np.array([cfo], np.float32)
    np.savez_compressed(os.path.join(HERE, "chirp_golden.npz"), **out)


def mcdpsk_got_chirp_vectors():
    """MC-DPSK behind an external chirp (SURVEY 8a row a16 with the Hilbert-FIR CFO correction): setChirpDetected(cfo) ->
    process -> getSoftBits on two 8-carrier frames: cfo 0.15 Hz (corrected, accepted) and cfo 12 Hz (corrected, rejected)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from projectultra_b200 import capi
    rng = np.random.default_rng(20261021)
    out = {}
    cfg = capi.mcdpsk_config(8, 2)
    for i, (snr, cfo) in enumerate(((12.0, 0.15), (15.0, 12.0))):
        tx = capi.mcdpsk_tx(cfg, capi.ldpc_encode(capi.R1_4, rng.integers(0, 256, 20, dtype=np.uint8)))
        p = float(np.mean(tx.astype(np.float64) ** 2))
        rx = (tx + rng.normal(0.0, np.sqrt(p / 10 ** (snr / 10)), len(tx))).astype(np.float32)
        llr, ready, after = R.mcdpsk_got_chirp(8, rx, cfo)
        out[f"m{i}_rx"] = rx
        out[f"m{i}_llr"] = llr
        out[f"m{i}_cfo"] = np.array([cfo, after], np.float32)
        out[f"m{i}_ready"] = np.array([int(ready)], np.int64)
    # the whole IWaveform sequence behind the dual chirp (MCDPSKWaveform detectSync -> setFrequencyOffset -> process): 10 dB with
    # +6.5 Hz CFO behind 800 samples of noise, and 4 dB without CFO
    sys.path.insert(0, os.path.dirname(HERE))
    from mcframes import mcdpsk_chirp_frame
    for i, (snr, lead, cfo) in enumerate(((10.0, 800, 6.5), (4.0, 0, 0.0))):
        rx = mcdpsk_chirp_frame(cfg, rng, snr, lead, cfo)
        llr, info, f, after = R.mcdpsk_chirp_receive(8, rx)
        out[f"r{i}_rx"] = rx
        out[f"r{i}_llr"] = llr
        out[f"r{i}_info"] = info.astype(np.int64)
        out[f"r{i}_f"] = f
        out[f"r{i}_after"] = np.array([after], np.float32)
    np.savez_compressed(os.path.join(HERE, "mcdpsk_chirp_golden.npz"), **out)


def frame_vectors():
    """Protocol-v2 multi-codeword frames: DataFrame::serialize -> encodeFrameWithLDPC -> +-6 LLRs with sign flips ->
    RxPipeline::decodeFrame, for a 1-codeword control frame, 2- and 5-codeword data frames (clean / a failing CW0 / a failing later
    codeword / too few codewords)."""
    sys.path.insert(0, os.path.dirname(HERE))
    import v2frames as V
    rng = np.random.default_rng(20261022)
    out = {}
    cases = [(R.R1_4, None, 0.0, 0, None), (R.R1_4, 30, 0.02, 0, None), (R.R1_2, 150, 0.0, 0, None), (R.R1_2, 150, 0.13, 0, None),
             (R.R3_4, 200, 0.01, 1, None), (R.R5_6, 64, 0.0, 0, None), (R.R1_2, 150, 0.2, 0, 2)]
    for i, (rate, plen, flip, drop, only_cw) in enumerate(cases):
        fr = R.control_frame_serialize() if plen is None else R.data_frame_serialize(rate, rng.integers(0, 256, plen, dtype=np.uint8))
        cws = R.frame_encode(rate, fr)
        llr = V.codeword_llrs(cws[:len(cws) - drop], rng, flip if only_cw is None else 0.0)
        if only_cw is not None:      # CW0 decodes, a later codeword does not
            llr[only_cw * 648:(only_cw + 1) * 648] = V.codeword_llrs(cws[only_cw:only_cw + 1], rng, flip)
        frame, info = R.frame_decode(rate, llr, len(cws) - drop)
        out[f"f{i}_rate"] = np.array([rate], np.int64)
        out[f"f{i}_frame"] = fr
        out[f"f{i}_codewords"] = cws
        out[f"f{i}_llr"] = llr.astype(np.float16).astype(np.float32)     # +-6 exactly representable: keeps the file small
        out[f"f{i}_ncw"] = np.array([len(cws) - drop], np.int64)
        out[f"f{i}_info"] = info.astype(np.int64)
        out[f"f{i}_out"] = frame
    out["count"] = np.array([len(cases)], np.int64)
    np.savez_compressed(os.path.join(HERE, "frame_golden.npz"), **out)


TRAINING_CFO_CASES = [("m1", R.DQPSK, R.R1_2, 40, 0.0, 20.0), ("m1", R.DQPSK, R.R1_2, 40, 12.0, 6.0), ("m1", R.QAM16, R.R1_2, 40, -7.5, 15.0),
                      ("m3", R.QAM32, R.R3_4, 60, 3.0, 18.0), ("m3", R.DQPSK, R.R3_4, 60, -7.5, 2.0), ("m1", R.D8PSK, R.R1_2, 40, 3.0, -20.0)]


def training_cfo_frame(i):
    preset, mod, rate, nbytes, tx_cfo, snr = TRAINING_CFO_CASES[i]
    rng = np.random.default_rng(4400 + i)
    cfg = (R.config_m1 if preset == "m1" else R.config_m3)(mod, rate)
    cfg.tx_cfo_hz = tx_cfo
    return cfg, rng.integers(0, 256, nbytes, dtype=np.uint8), rate, snr, rng


def training_cfo_vectors():
    """`reset(); processPresynced(span, 2)` WITHOUT setFrequencyOffset: the reference estimates the CFO from the two training symbols
    (estimateCFOFromTraining, src/ofdm/ofdm_sync.cpp:278-380).  Stored: the received frame, the CFO the reference adopts (as
    getFrequencyOffset() reports it for the differential no-pilot modes, whose tracker never moves it) and all soft bits."""
    out = {}
    for i in range(len(TRAINING_CFO_CASES)):
        cfg, data, rate, snr, rng = training_cfo_frame(i)
        rx = awgn(R.ofdm_tx(cfg, R.ldpc_encode(rate, data), 0), snr, rng)
        llr, snr_db, fc = R.ofdm_presynced(cfg, rx, 2, 0)
        out["c%d_cfg" % i] = np.frombuffer(bytes(cfg), dtype=np.uint8).copy()
        out["c%d_rx" % i] = rx
        out["c%d_llr" % i] = llr
        out["c%d_final_cfo" % i] = np.float32(fc)
    np.savez_compressed(os.path.join(HERE, "training_cfo_golden.npz"), **out)


if __name__ == "__main__":
    assert R.available(), "build oracle/_ref first: make -C oracle/ref_build"
    if "--only-training-cfo" in sys.argv:
        training_cfo_vectors()
        sys.exit(0)
    ldpc_vectors()
    ofdm_vectors()
    misc_vectors()
    psk_vectors()
    acquire_vectors()
    dpsk_acquire_vectors()
    chirp_vectors()
    mcdpsk_got_chirp_vectors()
    frame_vectors()
    training_cfo_vectors()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
