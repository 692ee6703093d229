"""Timing of the DPSK Barker acquisition path (SURVEY §8f next-2, config 4 frame: DQPSK R1/4, 139 392 samples): findPreamble +
demodulateSoft for B frames over AWGN.  python tools/dpsk_acquire_quick_bench.py [B]"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import refapi as R
from projectultra_b200 import capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ctx = capi.Context(0)
cfg = capi.dpsk_config(1, 384)
dem = capi.DpskDemodulator(ctx, cfg)
rng = np.random.default_rng(1)
pool = []
for _ in range(8):
    tx = capi.dpsk_tx(cfg, capi.ldpc_encode(capi.R1_4, rng.integers(0, 256, 20, dtype=np.uint8)), 0)
    pool.append((tx * (np.float32(0.5) / np.abs(tx).max())).astype(np.float32))
L = len(pool[0])
for snr in (10.0, -5.0, -14.0):
    tx = torch.from_numpy(np.stack([pool[i % 8] for i in range(B)])).cuda()
    p = (tx.double() ** 2).mean(dim=1, keepdim=True)
    g = torch.Generator(device="cuda"); g.manual_seed(int(snr) + 100)
    x = (tx + torch.randn(tx.shape, device="cuda", generator=g) * torch.sqrt(p / 10 ** (snr / 10)).float()).contiguous()
    del tx
    dem.receive_batch(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = dem.receive_batch(x); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("snr=%5.1f findPreamble+demodulateSoft B=%d L=%d ms=%.1f  %.1f kframes/s  %.1f GB/s of samples  found=%.3f" % (
        snr, B, L, ms, B / ms, B * L * 4 / ms / 1e6, (out[2] > 0).float().mean().item()), flush=True)
    if R.available():
        xs = x[:3].cpu().numpy()
        t0 = time.perf_counter()
        for f in xs: R.dpsk_receive(1, 384, f)
        print("   reference CPU (1 core): %.1f ms/frame" % ((time.perf_counter() - t0) / len(xs) * 1e3), flush=True)
    del x
