"""GPU parity of the protocol-v2 multi-codeword frame path (csrc/frame_v2.cu through the C ABI):
pu_frame_encode == v2::encodeFrameWithLDPC and pu_frame_decode_batch == RxPipeline::decodeFrame, byte for byte against the oracle
(pinned to the compiled reference and golden vectors in tests/test_oracle_fec.py) on ragged batches: clean frames, failing CW0,
failing later codewords, fewer codewords than TOTAL_CW, control frames, broken CRCs / magic, the legacy CW1 format."""
import numpy as np
import pytest

import oracleapi as O
import refapi as R
import v2frames as V

pytestmark = pytest.mark.gpu
RATES = [R.R1_4, R.R1_2, R.R2_3, R.R3_4, R.R5_6]


@pytest.fixture(scope="module")
def ctx():
    from projectultra_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def test_golden_v2_frames(ctx, golden):
    from projectultra_b200 import capi
    g = golden["frame"]
    for i in range(int(g["count"][0])):
        rate, ncw = int(g[f"f{i}_rate"][0]), int(g[f"f{i}_ncw"][0])
        cws = capi.frame_encode(rate, g[f"f{i}_frame"])
        assert cws.shape == g[f"f{i}_codewords"].shape and (cws == g[f"f{i}_codewords"]).all(), i
        dec = capi.LdpcDecoder(ctx, rate)
        frames, flen, info = dec.frame_decode_batch(g[f"f{i}_llr"].reshape(1, -1), ncw)
        want = g[f"f{i}_out"]
        assert (info[0] == g[f"f{i}_info"]).all() and int(flen[0]) == len(want) and (frames[0, :len(want)] == want).all(), (i, info[0])
        dec.close()


@pytest.mark.parametrize("rate", RATES)
def test_v2_frame_batch_matches_oracle(ctx, rate):
    import torch
    from projectultra_b200 import capi
    rng = np.random.default_rng(910 + rate)
    ncw = 6
    bpc = V.BYTES_PER_CW[rate]
    max_payload = bpc + (ncw - 1) * (bpc - 2) - 19
    rows, frames_tx = [], []
    def add(cws, flip=0.0, only=None, sigma=0.0):
        l = np.zeros(ncw * 648, np.float32)                  # unused codeword slots: erasures
        n = min(len(cws), ncw)
        l[:n * 648] = V.codeword_llrs(cws[:n], rng, 0.0 if only is not None else flip, sigma=sigma)
        if only is not None and only < n:
            l[only * 648:(only + 1) * 648] = V.codeword_llrs(cws[only:only + 1], rng, flip)
        rows.append(l)
    for plen in (0, 1, bpc - 19, bpc - 18, 3 * bpc, max_payload):
        fr = V.data_frame(rng.integers(0, 256, plen, dtype=np.uint8), rate, seq=plen)
        cws = capi.frame_encode(rate, fr)
        assert (cws == O.frame_encode(rate, fr)).all() and len(cws) == V.codewords_for(plen, rate) <= ncw
        frames_tx.append(fr)
        add(cws)
        add(cws, flip=0.04, sigma=1.0)
        add(cws, flip=0.25, only=0)
        if len(cws) > 1:
            add(cws, flip=0.25, only=len(cws) - 1)
    too_long = V.data_frame(rng.integers(0, 256, max_payload + 40, dtype=np.uint8), rate)     # TOTAL_CW > ncw: waiting
    add(capi.frame_encode(rate, too_long))
    c = V.control_frame()
    bad_c, bad_h, bad_m = c.copy(), frames_tx[4].copy(), frames_tx[4].copy()
    bad_c[19] ^= 1
    bad_h[16] ^= 0x40
    bad_m[1] = 0x4D
    for f in (c, bad_c, bad_h, bad_m, V.control_frame(ftype=0x21, seq=900), V.data_frame(b"connect", rate, ftype=0x12)):
        add(capi.frame_encode(rate, f))
    legacy = capi.frame_encode(rate, frames_tx[4])
    blk = rng.integers(0, 256, bpc, dtype=np.uint8)
    blk[0] = 0x11
    legacy[1] = O.ldpc_encode(rate, blk)
    add(legacy)
    rows.append(np.zeros(ncw * 648, np.float32))                                              # all erasures
    rows.append(rng.normal(0.0, 4.0, ncw * 648).astype(np.float32))                            # noise only
    llr = np.stack(rows)
    dec = capi.LdpcDecoder(ctx, rate)
    cap = ncw * bpc
    frames, flen, info = dec.frame_decode_batch(llr, ncw, frame_cap=cap)
    outcomes = set()
    for b in range(len(llr)):
        of, oi = O.frame_decode(rate, llr[b], ncw)
        assert (info[b] == oi).all(), (b, info[b], oi)
        assert int(flen[b]) == len(of) and (frames[b, :len(of)] == of).all(), b
        outcomes.add((int(oi[0]), int(oi[3]) > 0, int(oi[4]) > 0))
    assert {(1, False, True), (0, True, False), (0, True, True), (0, False, True), (0, False, False)} <= outcomes, outcomes
    for i, fr in enumerate(frames_tx):                                                         # clean frames come back as sent
        b = next(j for j in range(len(llr)) if int(info[j][0]) and int(flen[j]) == len(fr) and (frames[j, :len(fr)] == fr).all())
        assert b >= 0
    d = dec.frame_decode_batch(torch.from_numpy(llr).cuda(), ncw, frame_cap=cap)
    torch.cuda.synchronize()
    assert (d[1].cpu().numpy() == flen).all() and (d[2].cpu().numpy() == info).all()
    df = d[0].cpu().numpy()
    for b in range(len(llr)):
        assert (df[b, :int(flen[b])] == frames[b, :int(flen[b])]).all()
    dec.close()


def test_v2_frame_large_batch_property(ctx):
    """Size-independent property at scale: 16 384 five-codeword R1/2 frames (81 920 codewords in one LDPC launch), light noise: every
    frame comes back byte-identical to what was sent; frames with one codeword erased beyond repair report exactly one failure."""
    import torch
    from projectultra_b200 import capi
    rate, B, plen = R.R1_2, 16384, 150
    rng = np.random.default_rng(5)
    pool = [V.data_frame(rng.integers(0, 256, plen, dtype=np.uint8), rate, seq=i) for i in range(32)]
    cw = [capi.frame_encode(rate, f) for f in pool]
    ncw = len(cw[0])
    base = torch.from_numpy(np.stack([V.codeword_llrs(c, rng, 0.0, mag=1.0) for c in cw])).cuda()      # +-1
    idx = torch.arange(B, device="cuda") % 32
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    llr = (base[idx] * 4.0 + torch.randn((B, ncw * 648), device="cuda", generator=g) * 1.5).contiguous()
    broken = torch.arange(B, device="cuda") % 7 == 3
    llr[broken, 2 * 648:3 * 648] = torch.randn((int(broken.sum()), 648), device="cuda", generator=g) * 6.0
    dec = capi.LdpcDecoder(ctx, rate)
    frames, flen, info = dec.frame_decode_batch(llr, ncw)
    torch.cuda.synchronize()
    info, flen, frames, broken = info.cpu().numpy(), flen.cpu().numpy(), frames.cpu().numpy(), broken.cpu().numpy()
    want = np.stack(pool)[np.arange(B) % 32]
    good = ~broken
    assert (info[good, 0] == 1).all() and (info[good, 2] == ncw).all() and (flen[good] == want.shape[1]).all()
    assert (frames[good][:, :want.shape[1]] == want[good]).all()
    assert (info[broken, 0] == 0).all() and (info[broken, 3] == 1).all() and (info[broken, 2] == ncw - 1).all() and (flen[broken] == 0).all()
    dec.close()


def test_whole_frame_link_simulation_against_the_oracle(ctx):
    """linksim.FrameLinkSim: multi-codeword protocol-v2 frames through channel -> OFDM demod -> RxPipeline::decodeFrame on the GPU; every
    frame's verdict and bytes against the oracle (orc_ofdm_presynced + orc_frame_decode, both pinned to the compiled reference) on the
    same channel outputs; whole-frame FER falls with SNR and is never below the codeword-level failure it is made of."""
    import torch
    import oracleapi as O
    import refapi as R
    import v2frames
    from projectultra_b200 import capi, linksim
    for cfgargs, rate, chan, snrs, nbytes in (((48000, 1500, 512, 30, 1, 4, 2, 0, capi.DQPSK, capi.R1_2, 40.0, 0.0), capi.R1_2, "awgn", [0.0, 2.0, 6.0], 150),
                                              ((48000, 1500, 512, 30, 1, 4, 2, 1, capi.QAM16, capi.R3_4, 40.0, 0.0), capi.R3_4, "good", [8.0, 16.0, 30.0], 200)):
        cfg = capi.ModemConfig(*cfgargs)
        sim = linksim.FrameLinkSim(ctx, cfg, chan, payload_bytes=nbytes, pool=3, code_rate=rate)
        assert sim.n_cw == v2frames.codewords_for(nbytes, rate) and sim.n_cw >= 4
        assert (sim.frames_host[0] == v2frames.data_frame(sim.frames_host[0][17:17 + nbytes], rate, seq=1)).all()
        trials = 5
        si = np.repeat(np.arange(len(snrs)), trials)
        tr = np.tile(np.arange(trials), len(snrs))
        batch = sim.make_batch(snrs, si, tr)
        counters = torch.zeros((len(snrs), 6), dtype=torch.int64, device="cuda")
        rx, llr, frames, flen, info, good = sim.run_batch(batch, counters, keep=True)
        torch.cuda.synchronize()
        rx_h, info_h, flen_h, frames_h, good_h = rx.cpu().numpy(), info.cpu().numpy(), flen.cpu().numpy(), frames.cpu().numpy(), good.cpu().numpy()
        rcfg = R.ModemConfig.from_buffer_copy(bytes(cfg))
        for b in range(len(si)):
            soft, _, _ = O.ofdm_presynced(rcfg, rx_h[b], 2, 1, 0.0, 0.0)
            want_bytes, want_info = O.frame_decode(rate, soft[:sim.n_cw * 648], sim.n_cw)
            assert (info_h[b] == want_info).all(), (b, info_h[b], want_info)
            assert flen_h[b] == len(want_bytes) and (frames_h[b, :len(want_bytes)] == want_bytes).all(), b
            sent = sim.frames_host[batch["host"]["tx_index"][b]]
            assert bool(good_h[b]) == (want_info[0] == 1 and len(want_bytes) == len(sent) and (want_bytes == sent).all()), b
        c = counters.cpu().numpy()
        assert (c[:, 0] == trials).all() and (c[:, 3] == trials * sim.n_cw).all()
        assert c[0, 1] >= c[-1, 1] and (c[:, 1] * sim.n_cw >= c[:, 2]).all()
        if chan == "awgn":
            assert c[-1, 1] == 0
