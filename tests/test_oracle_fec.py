"""Pins the C oracle's FEC layer (oracle/pu_oracle_fec.c) against the reference's own KATs, the committed
golden vectors (generated from the unmodified reference by tests/golden/make_golden.py) and, when
oracle/_ref is built, against the reference itself on fresh random inputs.  CPU only."""
import numpy as np
import pytest

import oracleapi as O
import refapi as R

RATES = [R.R1_4, R.R1_2, R.R2_3, R.R3_4, R.R5_6]
# FNV-1a-64 of H_data rows per rate, extracted from the reference through encode(e_j)
FINGERPRINT = {R.R1_4: "b99341e1f8214bac", R.R1_2: "a458c7394a27efdc", R.R2_3: "14a59a7c36521cab",
               R.R3_4: "47f97faa839a0446", R.R5_6: "ea29e4d2c826be6e"}
EDGES = {R.R1_4: 2437, R.R1_2: 1623, R.R2_3: 1510, R.R3_4: 1134, R.R5_6: 756}


def fnv_rows(rows, k):
    h = 1469598103934665603
    for r in rows:
        for j in r:
            if j >= k:
                continue
            for b in int(j).to_bytes(4, "little"):
                h = ((h ^ b) * 1099511628211) % (1 << 64)
        h = ((h ^ 0xFF) * 1099511628211) % (1 << 64)
    return "%016x" % h


def test_mt19937_fisher_yates_kat():
    # tests/test_rng.cpp:24-39 documents 7 6 0 8 5 1 2 4 9 3 for mt19937(0x12345678) with rng() % i
    v = O.mt19937(0x12345678, 16)
    a = list(range(10))
    for n, i in enumerate(range(10, 1, -1)):
        j = int(v[n]) % i
        a[i - 1], a[j] = a[j], a[i - 1]
    assert a == [7, 6, 0, 8, 5, 1, 2, 4, 9, 3]
    # ISO C++ [rand.predef]: the 10000th output of a default-seeded mt19937 is 4123659995
    assert int(O.mt19937(5489, 10000)[-1]) == 4123659995


@pytest.mark.parametrize("rate", RATES)
def test_h_matrix_fingerprint(rate):
    k, m, rows = O.ldpc_build(rate)
    assert (k, m) == (R.RATE_K[rate], 648 - R.RATE_K[rate])  # tests/test_multiblock_ldpc.cpp:52-58
    assert sum(len(r) for r in rows) == EDGES[rate]
    assert fnv_rows(rows, k) == FINGERPRINT[rate]
    assert all(r[-1] == k + i for i, r in enumerate(rows))       # identity part, ldpc_decoder.cpp:124-128
    assert max(len(r) for r in rows) <= 7 and min(len(r) for r in rows) >= 2


def test_fallback_rates_use_r12_dimensions():
    # getCodeParams default branch (ldpc_decoder.cpp:33-34) -- but the RNG seed still uses the enum value
    for rate in (R.R1_3, R.R7_8):
        k, m, rows = O.ldpc_build(rate)
        assert (k, m) == (324, 324)


@pytest.mark.parametrize("rate", RATES)
def test_golden_ldpc(golden, rate):
    g = golden["ldpc"]
    assert (O.ldpc_encode(rate, g[f"r{rate}_data"][:R.RATE_K[rate] // 8]) == g[f"r{rate}_cw"]).all()
    info, ok, it = O.ldpc_decode_batch(rate, g[f"r{rate}_llr"])
    assert (info == g[f"r{rate}_info"]).all()
    assert (ok == g[f"r{rate}_ok"]).all()
    assert (it == g[f"r{rate}_iters"]).all()
    assert 0 < ok.sum() < len(ok)            # the vectors hold converging and non-converging codewords
    o, okm, itm = O.ldpc_decode_soft(rate, g[f"r{rate}_mb_llr"])
    assert (o == g[f"r{rate}_mb_out"]).all() and [int(okm), itm] == list(g[f"r{rate}_mb_ok"])


@pytest.mark.parametrize("rate", RATES)
@pytest.mark.parametrize("blocks", [1, 2, 5])
def test_encode_decode_identity(rate, blocks):
    # tests/test_multiblock_ldpc.cpp:104-230: encode -> +-6 LLR -> decodeSoft == data
    rng = np.random.default_rng(rate * 10 + blocks)
    k = R.RATE_K[rate]
    nbytes = (k // 8) * blocks
    data = rng.integers(0, 256, nbytes, dtype=np.uint8)
    cw = O.ldpc_encode(rate, data)
    nblk = -(-nbytes * 8 // k)
    assert len(cw) == 81 * nblk
    llr = np.where(np.unpackbits(cw) == 1, -6.0, 6.0).astype(np.float32)
    out, ok, it = O.ldpc_decode_soft(rate, llr)
    assert ok and it == 0
    assert (out[:nbytes] == data).all()


def test_decoder_edge_cases():
    # empty input -> {} and failure (ldpc_decoder.cpp:285-288); all-zero LLRs decode to the all-zero word
    out, ok, it = O.ldpc_decode_soft(R.R1_2, np.zeros(0, np.float32))
    assert len(out) == 0 and not ok
    out, ok, it = O.ldpc_decode_soft(R.R1_2, np.zeros(648, np.float32))
    assert ok and it == 0 and not out.any()
    # inverted LLRs of a valid codeword must not return the data (tests/test_comprehensive_modem.cpp:59-258)
    data = np.arange(40, dtype=np.uint8)
    cw = O.ldpc_encode(R.R1_2, data)
    llr = np.where(np.unpackbits(cw) == 1, 10.0, -10.0).astype(np.float32)
    out, ok, it = O.ldpc_decode_soft(R.R1_2, llr)
    assert not (out[:40] == data).all()
    # weak +-1.5 LLRs at R1/4 still decode
    cw = O.ldpc_encode(R.R1_4, data[:20])
    llr = np.where(np.unpackbits(cw) == 1, -1.5, 1.5).astype(np.float32)
    out, ok, it = O.ldpc_decode_soft(R.R1_4, llr)
    assert ok and (out[:20] == data[:20]).all()


def test_interleavers_golden(golden):
    g = golden["misc"]
    x = np.arange(648, dtype=np.float32)
    for bps, step in ((60, 181), (90, 271), (118, 355), (220, 325)):   # SURVEY a17 [probe]
        assert O.channel_interleaver_step(bps) == step
    for bps in (30, 60, 90, 118, 220, 708):
        y = O.channel_interleave(bps, x)
        assert (y == g[f"ci_{bps}"]).all()
        assert (O.channel_interleave(bps, y, inverse=True) == x).all()
    y = O.block_interleave(6, 108, x)
    assert (y == g["bi_6x108"]).all()
    assert (O.block_interleave(6, 108, y, inverse=True) == x).all()


@pytest.mark.ref
@pytest.mark.parametrize("rate", RATES)
def test_vs_reference_random(rate):
    rng = np.random.default_rng(100 + rate)
    k = R.RATE_K[rate]
    data = rng.integers(0, 256, 3 * 81, dtype=np.uint8)
    assert (R.ldpc_encode(rate, data) == O.ldpc_encode(rate, data)).all()
    cw = R.ldpc_encode(rate, data[:k // 8])
    bits = np.unpackbits(cw)[:648].astype(np.float32)
    for sigma in (0.45, 0.6, 0.75, 1.0, 1.3):
        y = (1 - 2 * bits)[None, :] + sigma * rng.standard_normal((24, 648)).astype(np.float32)
        llr = np.clip(2 * y / sigma ** 2, -10, 10).astype(np.float32)
        a, b = R.ldpc_decode_batch(rate, llr), O.ldpc_decode_batch(rate, llr)
        for u, v in zip(a, b):
            assert (u == v).all()
    mb = rng.standard_normal(648 * 2 + 123).astype(np.float32) * 4
    a, b = R.ldpc_decode_soft(rate, mb), O.ldpc_decode_soft(rate, mb)
    assert (a[0] == b[0]).all() and a[1:] == b[1:]
    a, b = R.ldpc_decode_soft(rate, mb[:77], max_iter=7), O.ldpc_decode_soft(rate, mb[:77], max_iter=7)
    assert (a[0] == b[0]).all() and a[1:] == b[1:]


# ---------------------------------------------------------------- protocol-v2 multi-codeword frames
def test_golden_v2_frames(golden):
    """encodeFrameWithLDPC and RxPipeline::decodeFrame restated in the oracle, against vectors produced by the unmodified reference."""
    g = golden["frame"]
    n = int(g["count"][0])
    for i in range(n):
        rate = int(g[f"f{i}_rate"][0])
        cws = O.frame_encode(rate, g[f"f{i}_frame"])
        assert cws.shape == g[f"f{i}_codewords"].shape and (cws == g[f"f{i}_codewords"]).all(), i
        out, info = O.frame_decode(rate, g[f"f{i}_llr"], int(g[f"f{i}_ncw"][0]))
        assert (info == g[f"f{i}_info"]).all() and len(out) == len(g[f"f{i}_out"]) and (out == g[f"f{i}_out"]).all(), (i, info)


@pytest.mark.ref
@pytest.mark.parametrize("rate", RATES)
def test_v2_frames_vs_reference(rate):
    if not R.available():
        pytest.skip("needs oracle/_ref (the compiled reference)")
    import v2frames as V
    rng = np.random.default_rng(77 + rate)
    seen = set()
    for plen in (0, 1, 5, 40, 200, 700):
        fr = R.data_frame_serialize(rate, rng.integers(0, 256, plen, dtype=np.uint8))
        mine = V.data_frame(fr[17:17 + plen], rate, seq=7, src_hash=int.from_bytes(bytes(fr[6:9]), "big"), dst_hash=int.from_bytes(bytes(fr[9:12]), "big"))
        assert len(mine) == len(fr) and (mine == fr).all(), plen          # the GPU tests' frame builder == DataFrame::serialize
        rc, oc = R.frame_encode(rate, fr), O.frame_encode(rate, fr)
        assert rc.shape == oc.shape and (rc == oc).all() and len(rc) == V.codewords_for(plen, rate), plen
        for flip in (0.0, 0.03, 0.12):
            l = V.codeword_llrs(rc, rng, flip)
            (rf, ri), (of, oi) = R.frame_decode(rate, l), O.frame_decode(rate, l)
            assert (ri == oi).all() and len(rf) == len(of) and (rf == of).all(), (plen, flip, ri, oi)
            seen.add((int(ri[0]), int(ri[3]) > 0, int(ri[4]) > 0))
            if flip == 0.0:
                assert ri[0] == 1 and (rf == fr).all()
        if len(rc) > 1:                                                    # fewer codewords than TOTAL_CW: waiting
            l = V.codeword_llrs(rc[:-1], rng)
            (rf, ri), (of, oi) = R.frame_decode(rate, l), O.frame_decode(rate, l)
            assert (ri == oi).all() and ri[0] == 0 and ri[4] == len(rc) and len(rf) == len(of) == 0
    assert (1, False, True) in seen and any(not s[0] for s in seen)
    # control frame, corrupted control CRC, corrupted header CRC, wrong magic, CW1 without its 0xD5 marker (legacy fallback)
    c = R.control_frame_serialize()
    assert (V.control_frame(src_hash=int.from_bytes(bytes(c[6:9]), "big"), dst_hash=int.from_bytes(bytes(c[9:12]), "big"),
                            seq=int.from_bytes(bytes(c[4:6]), "big"), payload=bytes(c[12:18]), flags=int(c[3])) == c).all()
    fr = V.data_frame(rng.integers(0, 256, 90, dtype=np.uint8), rate)
    bad_h, bad_m, bad_c = fr.copy(), fr.copy(), c.copy()
    bad_h[16] ^= 0x40
    bad_m[1] = 0x4D
    bad_c[19] ^= 1
    for f in (c, bad_c, bad_h, bad_m):
        l = V.codeword_llrs(O.frame_encode(rate, f), rng)
        (rf, ri), (of, oi) = R.frame_decode(rate, l), O.frame_decode(rate, l)
        assert (ri == oi).all() and len(rf) == len(of) and (rf == of).all(), (ri, oi)
    cws = O.frame_encode(rate, fr)
    bpc = V.BYTES_PER_CW[rate]
    legacy = np.zeros(bpc, np.uint8)
    legacy[:] = rng.integers(0, 256, bpc, dtype=np.uint8)
    legacy[0] = 0x11
    cws[1] = O.ldpc_encode(rate, legacy)
    l = V.codeword_llrs(cws, rng)
    (rf, ri), (of, oi) = R.frame_decode(rate, l), O.frame_decode(rate, l)
    assert (ri == oi).all() and ri[0] == 1 and len(rf) == len(of) and (rf == of).all()
