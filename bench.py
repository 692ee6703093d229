#!/usr/bin/env python
"""bench.py — decoded frames/s (OFDM demod + LDPC) of the batched Monte-Carlo receive chain on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # the reference's CPU path on the host cores

Workload (BASELINE.json: the metric's headline mode, "512-FFT DQPSK R1/2"): ModemConfig defaults M1 = 512-FFT, 30
carriers, CP 48, guard 4, DQPSK without pilots, LDPC R1/2, 40-byte payload; presynced frames of 2 LTS + 11 data
symbols = 7332 fp32 samples; AWGN at 13 SNR points -4..+8 dB (tool convention: mean frame power), i.e. a FER sweep
from "never converges" (50 iterations) through the waterfall to "error free".  A step is one pass of
demodulate -> soft demap -> LDPC decode -> frame/bit error counting over one batch of FRAMES_PER_POINT x 13 frames
per GPU.  The batch's channel outputs are produced by this repo's channel kernel BEFORE the timed region (inputs
resident in HBM) and are larger than L2, so no cache flush is needed between steps.  Ranks take disjoint trial
ranges (weak scaling, no data-path collective); the counter table is all-reduced once, inside the timed region.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decoded frames/sec (demod+LDPC)"
UNIT = "frames/s"
SNR_POINTS = [float(s) for s in range(-4, 9)]          # 13 points
PAYLOAD_BYTES = 40
POOL = 64
INFO_BITS = 324
FRAME_SAMPLES = 7332
N_LLR = 660
ALG_BYTES_DEMOD = 4 * FRAME_SAMPLES + 4 * 648           # per frame: samples read once + the codeword's LLRs written once
ALG_BYTES_LDPC = 4 * 648 + 41 + 5                       # per codeword: LLRs in, info bytes + ok + iters out


def workload_name(fpp):
    return ("M1 OFDM 512-FFT DQPSK R1/2 presynced (2 LTS + 11 data symbols, %d samples/frame), AWGN sweep %g..%g dB "
            "x %d frames/point/GPU, demod+demap+LDPC(flooding min-sum, <=50 it)+error count"
            % (FRAME_SAMPLES, SNR_POINTS[0], SNR_POINTS[-1], fpp))


def base_config(fpp, n):
    return {"workload": workload_name(fpp), "frames_per_step_per_gpu": fpp * len(SNR_POINTS),
            "snr_points_db": [SNR_POINTS[0], SNR_POINTS[-1], 1.0], "channel": "awgn", "payload_bytes": PAYLOAD_BYTES,
            "parallelism": "frames sharded over %d GPU(s), one counter all-reduce" % n}


# ------------------------------------------------------------------------------------------------ reference arm
_W = {}


def _ref_worker(args):
    """One forked worker: the reference's genie-timed recipe (processPresynced -> getSoftBits -> decodeSoft) over
    its share of the frames.  Returns (frames, syndrome-ok count)."""
    lo, hi = args
    import numpy as np
    x = _W["rx"][lo:hi]
    if len(x) == 0:
        return 0, 0
    t, info, ok = _W["api"].time_presynced_decode(_W["cfg"], x, 2)
    return len(x), int(ok.sum())


def run_reference(args):
    """CPU arm: oracle/_ref/libpu_ref.so (the unmodified reference compiled by oracle/ref_build/Makefile) when it is
    present, else the plain-C oracle port; one forked worker process per host core (avoids the
    reference's unsynchronised statics); frames regenerated on the CPU by the channel twin from the same
    (waveform, sigma, seed) triples the GPU arm uses."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import multiprocessing as mp
    import numpy as np
    import refapi as R
    import oracleapi as O
    import channelapi as CH
    from projectultra_b200.linksim import LinkSim, channel_preset

    kind = "reference" if R.available() else "port"
    api = R if kind == "reference" else O
    cores = len(os.sched_getaffinity(0))
    cfg = R.config_m1(R.DQPSK, R.R1_2)
    per_core = args.cpu_frames_per_core
    n_frames = min(max(len(SNR_POINTS), cores * per_core), args.frames_per_point * len(SNR_POINTS))
    fpp = (n_frames + len(SNR_POINTS) - 1) // len(SNR_POINTS)
    n_frames = fpp * len(SNR_POINTS)
    # frames in trial-major order so that every contiguous share holds the full SNR mix
    rng = np.random.default_rng(12345)
    payloads = rng.integers(0, 256, (POOL, PAYLOAD_BYTES), dtype=np.uint8)
    tx = [api.ofdm_tx(cfg, api.ldpc_encode(R.R1_2, p), 0) for p in payloads[:min(POOL, fpp)]]
    ch = channel_preset("awgn")
    rx = np.zeros((n_frames, FRAME_SAMPLES), np.float32)
    i = 0
    for t in range(fpp):
        w = tx[t % len(tx)]
        for s, snr in enumerate(SNR_POINTS):
            std = CH.noise_std(w, snr, 1)
            rx[i] = CH.channel_apply(ch, w, std, int(LinkSim.frame_seed(s, t)))
            i += 1
    _W.update(rx=rx, api=api, cfg=cfg)
    bounds = [(n_frames * c // cores, n_frames * (c + 1) // cores) for c in range(cores)]
    pool = mp.get_context("fork").Pool(cores)
    try:
        for _ in range(args.warmup):
            pool.map(_ref_worker, bounds)
        t0 = time.perf_counter()
        ok_total = 0
        for _ in range(args.steps):
            res = pool.map(_ref_worker, bounds)
            ok_total += sum(r[1] for r in res)
        dt = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    value = n_frames * args.steps / dt
    sample = ("%d frames/step (%d per SNR point, same 13-point sweep), %s on %d forked workers, -O3 x86-64 baseline"
              % (n_frames, fpp, "oracle/_ref (unmodified reference sources)" if kind == "reference" else "oracle C port",
                 cores))
    cfgd = base_config(args.frames_per_point, args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfgd, "info_bits_per_s": value * INFO_BITS, "gpu_launches": 0,
            "syndrome_ok_fraction": ok_total / (n_frames * args.steps),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, path):
        self.path, self.p, self.f = path, None, None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, gpus):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for ln in f:
                c = [v.strip() for v in ln.split(",")]
                if len(c) < 9:
                    continue
                try:
                    if int(c[0]) >= gpus:
                        continue
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                    power.append(float(c[3]))
                except ValueError:
                    continue
                for n, v in zip(names, c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


# ------------------------------------------------------------------------------------------------ this repo's arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from projectultra_b200 import capi, linksim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path is the product and there is no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU, bound to the CPUs next to that GPU: the end-to-end leg streams 1.5 GB of pinned host memory per
    # step and per GPU, and first-touch places those pages on the NUMA node of the thread that allocates them
    cpus_before = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local), (ncpu + 63) // 64)
        near = {c for c in range(ncpu) if (mask[c // 64] >> (c % 64)) & 1} & set(cpus_before)
        if near:
            os.sched_setaffinity(0, near)
    except Exception:   # noqa: BLE001  (no NVML / restricted container: keep the inherited affinity)
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = capi.Context(local)
    cfg = capi.ModemConfig(48000, 1500, 512, 30, 1, 4, 2, 0, capi.DQPSK, capi.R1_2, 40.0, 0.0)
    sim = linksim.LinkSim(ctx, cfg, "awgn", payload_bytes=PAYLOAD_BYTES, pool=POOL)
    assert sim.L == FRAME_SAMPLES
    fpp = args.frames_per_point
    n_snr = len(SNR_POINTS)
    B = fpp * n_snr
    # this rank's frames: trials [rank*fpp, (rank+1)*fpp) of every SNR point, trial-major so SNR points interleave
    trials = np.repeat(np.arange(rank * fpp, (rank + 1) * fpp, dtype=np.int64), n_snr)
    si = np.tile(np.arange(n_snr, dtype=np.int64), fpp)
    batch = sim.make_batch(SNR_POINTS, si, trials)
    rx = linksim.channel_apply(ctx, sim.ch, sim.tx_pool, batch["tx_index"], batch["noise_std"], batch["seed"])
    llr = torch.zeros((B, 648), dtype=torch.float32, device=dev)
    info = torch.empty((B, sim.ldpc.info_bytes), dtype=torch.uint8, device=dev)
    ok = torch.empty(B, dtype=torch.uint8, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    counters = torch.zeros((n_snr, 6), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def step(ev=None):
        if ev is not None:
            ev[0].record()
        sim.ofdm.presynced_batch(rx, 2, llr=llr, want_aux=False)
        if ev is not None:
            ev[1].record()
        sim.ldpc.decode_batch(llr, info, ok, iters)
        if ev is not None:
            ev[2].record()
        linksim.count_errors(ctx, info, ok, iters, sim.payload_pool, batch["tx_index"], batch["bins"], PAYLOAD_BYTES, counters)
        if ev is not None:
            ev[3].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    counters.zero_()
    barrier()
    sampler = ClockSampler(os.path.join(ROOT, "gpurun_out", "bench_clocks.csv")) if rank == 0 else None
    if sampler:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ctx.kernel_launches
    barrier()
    e0.record()
    for k in range(args.steps):
        step(evs[k])
    linksim.allreduce_counters(counters)          # the path's only collective: once per sweep
    e1.record()
    barrier()
    launches = ctx.kernel_launches - n0
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_demod = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    ms_ldpc = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    ms_count = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    c = counters.cpu().numpy()

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D of the samples, D2H of the decoded bytes/flags
    rx_host = torch.empty((B, FRAME_SAMPLES), dtype=torch.float32, pin_memory=True)
    rx_host.copy_(rx)
    torch.cuda.synchronize()
    rx_np = rx_host.numpy()
    info_h = np.zeros((B, sim.ldpc.info_bytes), np.uint8)
    ok_h = np.zeros(B, np.uint8)
    it_h = np.zeros(B, np.int32)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(args.warmup, 2)):
        linksim.receive_decode(sim.ofdm, sim.ldpc, rx_np, info=info_h, ok=ok_h, iters=it_h)
    barrier()
    tb0 = ctx.transfer_bytes
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        linksim.receive_decode(sim.ofdm, sim.ldpc, rx_np, info=info_h, ok=ok_h, iters=it_h)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tb1 = ctx.transfer_bytes       # bytes the library actually moved (only the FFT windows of the symbols the kernel reads cross PCIe)
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(t.item())
    e2e_matches = bool((info_h == info.cpu().numpy()).all() and (ok_h == ok.cpu().numpy()).all())
    barrier()
    clocks = sampler.stop(world) if sampler else None

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            with open(peaks_path) as f:
                peak, peak_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        else:
            peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("ofdm_presynced_kernel_bytes_per_frame"):
                traffic = tj["ofdm_presynced_kernel_bytes_per_frame"] * B
        ach = ALG_BYTES_DEMOD * B / (ms_demod * 1e-3) / 1e9
        demod_kernel = sim.ofdm.last_kernel
        iters_run = float((iters.float() + ok.float()).clamp(max=50).mean().item())
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(base_config(fpp, world), l2="inputs larger than L2: %.0f MB of samples per step per GPU, no flush"
                               % (B * FRAME_SAMPLES * 4 / 1e6)),
                "info_bits_per_s": value * INFO_BITS, "gpu_launches": int(launches),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * (tb1[0] - tb0[0]) // e2e_steps,
                        "d2h_bytes_per_step": world * (tb1[1] - tb0[1]) // e2e_steps, "steps": e2e_steps,
                        "host_buffer_bytes_per_step": world * B * FRAME_SAMPLES * 4,
                        "api": "pu_receive_decode_batch(PU_MEM_HOST)", "matches_device_path": e2e_matches},
                "roofline": {"kernel": demod_kernel, "bound": "hbm", "achieved": ach, "peak": peak,
                             "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_frame": ALG_BYTES_DEMOD, "ms_per_launch": ms_demod},
                "stages_ms": {demod_kernel: ms_demod, "ldpc_flood_kernel": ms_ldpc, "count_errors_kernel": ms_count},
                "ldpc": {"codewords_per_s": B / (ms_ldpc * 1e-3), "avg_iterations_run": iters_run,
                         "edge_updates_per_s": 2 * 1623 * iters_run * B / (ms_ldpc * 1e-3),
                         "hbm_gbs": ALG_BYTES_LDPC * B / (ms_ldpc * 1e-3) / 1e9},
                "fer": [round(float(r[1]) / max(int(r[0]), 1), 5) for r in c],
                "frames_counted": int(c[:, 0].sum())}
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, cpus_before)      # the reference arm uses every host core
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def cpu_baseline(args):
    """The reference arm on a bounded sample, run in a child process (fork workers must not inherit a CUDA context)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(args.cpu_steps), "--warmup", "1",
           "--cpu-frames-per-core", str(args.cpu_frames_per_core)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: " + r.stderr[-300:]}
    except Exception as e:   # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-point", type=int, default=4096, help="frames per SNR point per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-frames-per-core", type=int, default=1040, help="reference arm: frames per core per step")
    ap.add_argument("--cpu-steps", type=int, default=10, help="steps of the cpu_baseline leg of the default run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
