"""Timing of the dual-chirp receive path: OFDM_CHIRP frames (57 600-sample dual chirp + M1 DQPSK R1/2 body)
over AWGN: detectDualChirp + processPresynced for B frames; with `mcdpsk` the same for MC-DPSK frames (dual chirp + 8-carrier DQPSK
R1/4 body, +3 Hz CFO: detectDualChirp + Hilbert-FIR correction + processGotChirp).  python tools/chirp_quick_bench.py [B] [mcdpsk]"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import refapi as R, oracleapi as O
from projectultra_b200 import capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ctx = capi.Context(0)
if len(sys.argv) > 2 and sys.argv[2] == "mcdpsk":
    from mcframes import mcdpsk_chirp_frame
    mcfg = capi.mcdpsk_config(8, 2)
    dem = capi.McDpskDemodulator(ctx, mcfg)
    rng = np.random.default_rng(2)
    for snr, cfo in ((12.0, 3.0), (0.0, 0.0)):
        pool = [mcdpsk_chirp_frame(mcfg, rng, snr, 500, cfo, tail=500) for _ in range(16)]
        x = torch.from_numpy(np.stack([pool[i % 16] for i in range(B)])).cuda().contiguous()
        L = x.shape[1]
        dem.chirp_receive_batch(x, llr_stride=648); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        capi.chirp_search_stats()
        e0.record(); out = dem.chirp_receive_batch(x, llr_stride=648); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        st = capi.chirp_search_stats()
        print("snr=%5.1f cfo=%.1f MC-DPSK detectDualChirp+process B=%d L=%d ms=%.1f  %.1f kframes/s  detected=%.3f  with soft bits=%.3f  searches=%d verify rounds=%d fine runs=%d" % (
            snr, cfo, B, L, ms, B / ms, out[2][:, 0].float().mean().item(), (out[1] > 0).float().mean().item(), st[0], st[1], st[2]), flush=True)
        if R.available():
            xs = x[:2].cpu().numpy()
            t0 = time.perf_counter()
            for f in xs: R.mcdpsk_chirp_receive(8, f)
            print("   reference CPU (1 core): %.1f ms/frame" % ((time.perf_counter() - t0) / len(xs) * 1e3), flush=True)
        del x
    sys.exit(0)
cfg = R.config_m1(R.DQPSK, R.R1_2)
dem = capi.OfdmDemodulator(ctx, capi.ModemConfig.from_buffer_copy(bytes(cfg)))
rng = np.random.default_rng(1)
chirp = capi.chirp_generate()
pool = [np.concatenate([np.zeros(500, np.float32), chirp, O.ofdm_tx(cfg, O.ldpc_encode(R.R1_2, rng.integers(0, 256, 40, dtype=np.uint8)), 0),
                        np.zeros(500, np.float32)]) for _ in range(8)]
L = len(pool[0])
for snr in (15.0, -5.0):
    tx = torch.from_numpy(np.stack([pool[i % 8] for i in range(B)])).cuda()
    g = torch.Generator(device="cuda"); g.manual_seed(int(snr) + 50)
    x = (tx + torch.randn(tx.shape, device="cuda", generator=g) * float(np.sqrt(0.02 / 10 ** (snr / 10)))).contiguous()
    del tx
    dem.chirp_receive_batch(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    capi.chirp_search_stats(); capi.chirp_phase_cycles()
    e0.record(); out = dem.chirp_receive_batch(x); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = capi.chirp_search_stats()
    print("snr=%5.1f detectDualChirp+process B=%d L=%d ms=%.1f  %.1f kframes/s  detected=%.3f  searches=%d verify rounds=%d fine runs=%d" % (
        snr, B, L, ms, B / ms, out[2][:, 0].float().mean().item(), st[0], st[1], st[2]), flush=True)
    cyc = capi.chirp_phase_cycles()
    print("   cycles per search: decimate %d, estimates %d, rank %d, exact coarse %d, fine ranking %d, exact fine %d" % tuple(c // max(st[0], 1) for c in cyc[:6]))
    if R.available():
        xs = x[:2].cpu().numpy()
        t0 = time.perf_counter()
        for f in xs: R.ofdm_chirp_receive(cfg, f)
        print("   reference CPU (1 core): %.1f ms/frame" % ((time.perf_counter() - t0) / len(xs) * 1e3), flush=True)
    del x
