"""Timing of the protocol-v2 multi-codeword frame path: B data frames of 150 payload bytes = 5 R1/2 codewords each,
BPSK-over-AWGN LLRs at two noise levels, through pu_frame_decode_batch (one LDPC launch over 5 B codewords + the assembly kernel).
python tools/frame_quick_bench.py [B]"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import refapi as R, v2frames as V
from projectultra_b200 import capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rate, plen = capi.R1_2, 150
ctx = capi.Context(0)
dec = capi.LdpcDecoder(ctx, rate)
rng = np.random.default_rng(3)
pool = [V.data_frame(rng.integers(0, 256, plen, dtype=np.uint8), rate, seq=i) for i in range(64)]
cw = [capi.frame_encode(rate, f) for f in pool]
ncw = len(cw[0])
base = torch.from_numpy(np.stack([V.codeword_llrs(c, rng, 0.0, mag=1.0) for c in cw])).cuda()
idx = torch.arange(B, device="cuda") % 64
g = torch.Generator(device="cuda"); g.manual_seed(1)
for sigma in (0.5, 0.74):
    y = base[idx] + sigma * torch.randn((B, ncw * 648), device="cuda", generator=g)
    llr = (2.0 * y / sigma ** 2).clamp(-10, 10).contiguous()
    del y
    dec.frame_decode_batch(llr, ncw); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); frames, flen, info = dec.frame_decode_batch(llr, ncw); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("sigma=%.2f v2 frames (%d codewords each) B=%d ms=%.2f  %.2f Mframes/s  %.2f Mcodewords/s  frame success=%.4f  cw failed/frame=%.3f" % (
        sigma, ncw, B, ms, B / ms / 1e3, B * ncw / ms / 1e3, info[:, 0].float().mean().item(), info[:, 3].float().mean().item()), flush=True)
    if R.available():
        xs = llr[:64].cpu().numpy()
        t0 = time.perf_counter()
        for f in xs: R.frame_decode(rate, f, ncw)
        print("   reference CPU (1 core): %.3f ms/frame" % ((time.perf_counter() - t0) / len(xs) * 1e3), flush=True)
    del llr
