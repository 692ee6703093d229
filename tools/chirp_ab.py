"""A/B of the two-tier chirp search against the brute-force kernel on Watterson-channel frames (the FER tables of two builds differed
in ~2e-4 of the OFDM_CHIRP frames): python tools/chirp_ab.py [frames] [snr_db] [channel]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import refapi as R
from projectultra_b200 import capi, linksim

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
snr = float(sys.argv[2]) if len(sys.argv) > 2 else 12.0
chan = sys.argv[3] if len(sys.argv) > 3 else "good"
ctx = capi.Context(0)
cfg = capi.ModemConfig.from_buffer_copy(bytes(R.config_m1(R.DQPSK, R.R1_2)))
sim = linksim.LinkSim(ctx, cfg, chan, payload_bytes=40, pool=32, layout="chirp", peak=0.5, precision="fast")
# silence behind the frame as the sweep tables carry it
pool = torch.cat([sim.tx_pool, torch.zeros(sim.tx_pool.shape[0], 2400, device=sim.tx_pool.device)], dim=1).contiguous()
std = None
dev = pool.device
idx = (torch.arange(B, device=dev) % pool.shape[0]).to(torch.int32)
if std is None:
    p = (pool.double() ** 2).mean(dim=1)
    std_pool = (p.sqrt() * 10 ** (-snr / 20)).float()
    std = std_pool[idx.long()]
seed = torch.arange(B, device=dev, dtype=torch.int64) + 777
rx = linksim.channel_apply(ctx, sim.ch, pool, idx, std, seed, None)
dem = sim.ofdm
os.environ.pop("PU_CHIRP_SEARCH", None)
for guard in (None, "8", "32"):
    if guard: os.environ["PU_CHIRP_GUARD"] = guard
    else: os.environ.pop("PU_CHIRP_GUARD", None)
    capi.chirp_search_stats()
    fast = dem.chirp_receive_batch(rx, llr_stride=648)
    st = capi.chirp_search_stats()
    if guard is None:
        os.environ["PU_CHIRP_SEARCH"] = "exact"
        slow = dem.chirp_receive_batch(rx, llr_stride=648)
        os.environ.pop("PU_CHIRP_SEARCH", None)
        torch.cuda.synchronize()
    fi, si = fast[2].cpu().numpy(), slow[2].cpu().numpy()
    fv, sv = fast[3].cpu().numpy().view(np.uint32), slow[3].cpu().numpy().view(np.uint32)
    bad = np.flatnonzero((fi != si).any(axis=1) | (fv != sv).any(axis=1))
    print("guard %s: %d frames, %d detected, %d differ from the brute-force search; searches %d, coarse rounds %d, fine runs %d" % (
        guard or "4 (default)", B, int((si[:, 0] != 0).sum()), len(bad), st[0], st[1], st[2]))
    for b in bad[:6]:
        print("   frame %d: two-tier %s %s | brute force %s %s" % (b, fi[b].tolist(), fast[3][b].cpu().numpy().tolist(), si[b].tolist(), slow[3][b].cpu().numpy().tolist()))
