"""Timing of the acquisition path on the config-1 frame: M1 512-FFT DQPSK R1/2, generatePreamble() +
modulate() = 10124 samples, AWGN, process() in 960-sample chunks (tools/test_mode_snr.cpp).  python tools/acquire_quick_bench.py [B]"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracleapi as O, refapi as R
from projectultra_b200 import capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cfg = R.config_m1(R.DQPSK, R.R1_2)
ctx = capi.Context(0)
dem = capi.OfdmDemodulator(ctx, capi.ModemConfig.from_buffer_copy(bytes(cfg)))
rng = np.random.default_rng(1)
pool = [O.ofdm_tx(cfg, O.ldpc_encode(R.R1_2, rng.integers(0, 256, 40, dtype=np.uint8)), 1) for _ in range(16)]
L = len(pool[0])
for snr in (25.0, 17.0, 10.0):
    tx = torch.from_numpy(np.stack([pool[i % 16] for i in range(B)])).cuda()
    p = (tx.double() ** 2).mean(dim=1, keepdim=True)
    g = torch.Generator(device="cuda"); g.manual_seed(int(snr))
    x = (tx + torch.randn(tx.shape, device="cuda", generator=g) * torch.sqrt(p / 10 ** (snr / 10)).float()).contiguous()
    for name, fn in (("acquire", lambda: dem.acquire_batch(x)), ("process", lambda: dem.process_batch(x))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        info = out[0] if name == "acquire" else out[2]
        print("snr=%4.1f %-8s B=%d ms=%.2f  %.3f Mframes/s  synced=%.3f  calls_avg=%.2f" % (
            snr, name, B, ms, B / ms / 1e3, info[:, 0].float().mean().item(), info[:, 3].float().mean().item()), flush=True)
    if R.available():
        xs = x[:8].cpu().numpy()
        t0 = time.perf_counter()
        for f in xs: R.ofdm_process_info(cfg, f, 960)
        print("   reference CPU (1 core): %.2f ms/frame" % ((time.perf_counter() - t0) / len(xs) * 1e3), flush=True)
