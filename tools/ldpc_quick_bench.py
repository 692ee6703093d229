"""Quick LDPC-only timing (config 2 shape) used during development: python tools/ldpc_quick_bench.py [B]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import oracleapi as O
from projectultra_b200 import capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
ctx = capi.Context(0)
EDGES = {0: 2437, 2: 1623, 3: 1510, 4: 1134, 5: 756}
K = {0: 162, 2: 324, 3: 432, 4: 486, 5: 540}
SIG = {0: (0.7, 1.12, 1.5), 2: (0.5, 0.71, 0.95), 3: (0.45, 0.61, 0.8), 4: (0.4, 0.57, 0.8), 5: (0.4, 0.58, 0.8)}
for rate in (0, 2, 3, 4, 5):
    dec = capi.LdpcDecoder(ctx, rate)
    rng = np.random.default_rng(rate)
    P = 4096
    cws = np.stack([np.unpackbits(O.ldpc_encode(rate, rng.integers(0, 256, K[rate] // 8, dtype=np.uint8)))[:648] for _ in range(64)])
    bits = torch.from_numpy(cws[np.arange(P) % 64].astype(np.float32)).cuda()
    for name, sigma in zip(("easy", "waterfall", "stress"), SIG[rate]):
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        y = (1 - 2 * bits)[torch.arange(B, device="cuda") % P] + sigma * torch.randn((B, 648), device="cuda", generator=g)
        llr = torch.clamp(2 * y / sigma ** 2, -10, 10).contiguous()
        del y
        info = torch.empty((B, dec.info_bytes), dtype=torch.uint8, device="cuda")
        ok = torch.empty(B, dtype=torch.uint8, device="cuda"); it = torch.empty(B, dtype=torch.int32, device="cuda")
        for _ in range(2): dec.decode_batch(llr, info, ok, it)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); n = 3
        for _ in range(n): dec.decode_batch(llr, info, ok, it)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        iters_run = (it.float() + ok.float()).clamp(max=50).mean().item()
        upd = 2 * EDGES[rate] * iters_run * B / (ms * 1e-3)
        print(f"rate={rate} {name:9s} sigma={sigma} ok={ok.float().mean().item():.3f} it_avg={it.float().mean().item():.2f} "
              f"ms={ms:.2f} cw/s={B/(ms*1e-3):.3e} edge_upd/s={upd:.3e} GB/s={B*(2592+dec.info_bytes+5)/(ms*1e-3)/1e9:.1f}", flush=True)
        del llr
