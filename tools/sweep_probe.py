import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from projectultra_b200 import capi
ctx = capi.Context(0)
cfg = capi.ModemConfig(48000, 1500, 512, 30, 1, 4, 2, 0, capi.DQPSK, capi.R1_2, 40.0, 0.0)
for ch in ("awgn", "good"):
    for block in (4096, 53248):
        mode = capi.sweep_mode(capi.WF_OFDM, cfg, capi.R1_2, 40, ch, -4, 1, 13, precision="fast")
        capi.Sweep([mode], trials_per_point=4096, block_trials=4096).run(ctx)
        sw = capi.Sweep([mode], trials_per_point=4096 * 100, block_trials=block, batch_bytes=53248 * 7332 * 4)
        c, st = sw.run(ctx)
        print(ch, block, "frames %d s %.3f setup %.3f wait %.3f fill %.3f enq %.3f -> %.2f Mf/s (excl. setup %.2f)" % (
            st.frames_run, st.seconds, st.setup_seconds, st.wait_seconds, st.fill_seconds, st.enqueue_seconds, st.frames_run / st.seconds / 1e6,
            st.frames_run / (st.seconds - st.setup_seconds) / 1e6), flush=True)
