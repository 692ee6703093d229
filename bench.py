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
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
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
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        self.p = None
        return self.window(None, None, gpus)

    def window(self, t0, t1, gpus):
        """Median SM clock / throttle reasons of the samples taken between wall-clock times t0 and t1 (None = all)."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.f is not None and not self.f.closed:
            self.f.flush()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            lines = open(self.path).read().splitlines()
        except OSError:
            return out
        for ln in lines:
            c = [v.strip() for v in ln.split(",")]
            if len(c) < 10:
                continue
            try:
                if t0 is not None:
                    ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if ts < t0 or ts > t1:
                        continue
                if int(c[1]) >= gpus:
                    continue
                sm.append(float(c[2]))
                mx.append(float(c[3]))
                power.append(float(c[4]))
            except ValueError:
                continue
            for n, v in zip(names, c[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


# ------------------------------------------------------------------------------------------------ this repo's arm
# BASELINE.json configs behind --workload.  "m1" is the headline the metric is quoted on (the driver's default run); the others
# are the same contract on configs 2-4 so that their numbers come from this file and not from ad-hoc tools.
WORKLOADS = {
    "m1": dict(kind="ofdm", cfg=(48000, 1500, 512, 30, 1, 4, 2, 0, "DQPSK", "R1_2", 40.0, 0.0), rate="R1_2", payload=40,
               channel="awgn", snr=[float(s) for s in range(-4, 9)], info_bits=324,
               name="M1 OFDM 512-FFT DQPSK R1/2 presynced (2 LTS + 11 data symbols, 7332 samples/frame), AWGN sweep -4..8 dB"),
    "m3": dict(kind="ofdm", cfg=(48000, 1500, 1024, 59, 1, 0, 4, 1, "QAM32", "R3_4", 40.0, 0.0), rate="R3_4", payload=60,
               channel="good", snr=[float(s) for s in range(8, 21)], info_bits=486,
               name="config 3: M3 OFDM 1024-FFT NVIS 32QAM R3/4 pilots/4 presynced, Watterson 'good' channel seeds, sweep 8..20 dB"),
    "dpsk": dict(kind="dpsk", rate="R1_4", payload=20, channel="poor", snr=[float(s) for s in range(-11, 18, 2)], info_bits=162,
                 name="config 4: single-carrier DQPSK 125 baud R1/4 (Barker preamble, genie data start), Watterson 'poor' channel, "
                      "FER-vs-SNR sweep -11..17 dB"),
}


class Timer:
    """CUDA-event stopwatch on torch's current stream (the stream every C-ABI call of this file launches on)."""

    def __init__(self, torch):
        self.t = torch
        self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        self.a.record()
        return self

    def __exit__(self, *exc):
        self.b.record()

    def ms(self):
        self.t.cuda.synchronize()
        return self.a.elapsed_time(self.b)


def measured_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    return 6650.0, "fallback of B200_PROFILING.md"


def traffic_of(kernel):
    """Per-frame DRAM bytes of `kernel` from the committed ncu --set full capture (profiles/traffic.json: dram__bytes_read.sum +
    dram__bytes_write.sum of one launch / frames of that launch).  A constant of the last profiled build, not of this run."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None, None
    with open(tpath) as f:
        tj = json.load(f)
    e = tj.get("kernels", {}).get(kernel)
    return (e["dram_bytes_per_unit"], e.get("source")) if e else (None, None)


def bind_near_cpus(local):
    """One process per GPU, bound to the CPUs NVML reports next to that GPU (the end-to-end leg streams GBs of pinned host memory
    per step and first-touch places those pages on the allocating thread's NUMA node).  When every GPU reports the same CPU set
    (virtualised hosts: the driver's 8-GPU box shows 0-31 / NUMA 0 for all) split that set evenly between the local ranks instead,
    so the ranks' staging threads at least do not share cores."""
    before = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        ncpu = os.cpu_count() or 1
        n_local = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
        sets = []
        for g in range(max(n_local, local + 1)):
            mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(g), (ncpu + 63) // 64)
            sets.append(frozenset(c for c in range(ncpu) if (mask[c // 64] >> (c % 64)) & 1) & frozenset(before))
        near = sets[local]
        if near and n_local > 1 and all(x == near for x in sets[:n_local]):
            cpus = sorted(near)
            per = max(1, len(cpus) // n_local)
            near = frozenset(cpus[local * per:(local + 1) * per]) or near
        if near:
            os.sched_setaffinity(0, near)
    except Exception:   # noqa: BLE001  (no NVML / restricted container: keep the inherited affinity)
        pass
    return before


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
    cpus_before = bind_near_cpus(local)
    if world > 1:
        # stdout carries exactly one JSON line: whatever NCCL prints while the communicator comes up (its "NCCL version ..." banner when
        # the box exports NCCL_DEBUG) is sent to stderr by pointing fd 1 there for the duration of the first collective
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    ctx = capi.Context(local)
    if args.workload == "ldpc":
        return run_ldpc(args, ctx, dev, world, rank)
    wl = WORKLOADS[args.workload]
    snr_points = wl["snr"]
    rate = getattr(capi, wl["rate"])
    if wl["kind"] == "ofdm":
        c = list(wl["cfg"])
        c[8], c[9] = getattr(capi, c[8]), getattr(capi, c[9])
        cfg = capi.ModemConfig(*c)
        sim = linksim.LinkSim(ctx, cfg, wl["channel"], payload_bytes=wl["payload"], pool=POOL, code_rate=rate, precision=args.precision)
    else:
        cfg = capi.dpsk_config(1, 384)       # DQPSK, 125 baud at 48 kHz (DPSKConfig defaults, tools/test_dpsk_snr.cpp:28-52)
        sim = linksim.LinkSim(ctx, cfg, wl["channel"], payload_bytes=wl["payload"], pool=16, code_rate=rate, peak=0.5)
    if args.workload == "m1":
        assert sim.L == FRAME_SAMPLES
    L = sim.L
    fpp = args.frames_per_point if args.workload == "m1" else max(1, args.frames_per_point * FRAME_SAMPLES // L)
    n_snr = len(snr_points)
    B = fpp * n_snr
    pb = wl["payload"]
    # this rank's frames: trials [rank*fpp, (rank+1)*fpp) of every SNR point, trial-major so SNR points interleave
    trials = np.repeat(np.arange(rank * fpp, (rank + 1) * fpp, dtype=np.int64), n_snr)
    si = np.tile(np.arange(n_snr, dtype=np.int64), fpp)
    batch = sim.make_batch(snr_points, si, trials)
    rx = sim.make_rx(batch)
    bufs = dict(llr=torch.zeros((B, 648), dtype=torch.float32, device=dev),
                info=torch.empty((B, sim.ldpc.info_bytes), dtype=torch.uint8, device=dev),
                ok=torch.empty(B, dtype=torch.uint8, device=dev), iters=torch.empty(B, dtype=torch.int32, device=dev))
    counters = torch.zeros((n_snr, 6), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def step(ev=None):
        sim.receive_count(batch, rx, counters, ev=ev, bufs=bufs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    counters.zero_()
    barrier()
    sampler = ClockSampler(os.path.join(ROOT, "gpurun_out", "bench_clocks.csv")) if rank == 0 else None
    if sampler:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        sampler.start()
    # ---- the contract's timed region: EXACTLY --steps steps between two barriers, CUDA events, max over ranks
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    n0 = ctx.kernel_launches
    barrier()
    with Timer(torch) as tm:
        for k in range(args.steps):
            step(evs[k])
        linksim.allreduce_counters(counters)          # the path's only collective: once per sweep
    barrier()
    launches = ctx.kernel_launches - n0
    ms_total = max_over_ranks(tm.ms())
    ms_demod = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    ms_ldpc = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    ms_count = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    c = counters.cpu().numpy()
    demod_kernel = sim.demod.last_kernel if hasattr(sim.demod, "last_kernel") else wl["kind"] + "_demod_kernels"
    iters_run = float((bufs["iters"].float() + bufs["ok"].float()).clamp(max=50).mean().item())

    # ---- sustained: the same step back to back for >= --sustain-seconds, so that the clocks the part settles at under a long
    #      run are sampled (the K-step region above is a few tens of ms: a burst)
    sustained = None
    if args.sustain_seconds > 0:
        n_sus = max(args.steps, int(args.sustain_seconds * 1e3 / max(ms_total / args.steps, 1e-3)) + 1)
        barrier()
        mark = time.time()
        with Timer(torch) as ts:
            for _ in range(n_sus):
                step()
        ms_sus = max_over_ranks(ts.ms())
        barrier()
        sustained = {"value": world * B * n_sus / (ms_sus * 1e-3), "unit": UNIT, "steps": n_sus, "seconds": ms_sus * 1e-3,
                     "ms_per_step": ms_sus / n_sus, "window": [mark, time.time()]}

    # ---- the other precision on the same inputs (OFDM modes with an FMA form): demodulator time + the whole step's counters
    other = None
    if wl["kind"] == "ofdm" and demod_kernel in ("ofdm_fast512_kernel", "ofdm_diff512_kernel"):
        alt = "exact" if sim.ofdm.precision == "fast" else "fast"
        sim.ofdm.set_precision(alt)
        c_alt = torch.zeros_like(counters)
        sim.receive_count(batch, rx, c_alt, bufs=bufs)
        with Timer(torch) as ta:
            for _ in range(5):
                sim.ofdm.presynced_batch(rx, 2, llr_stride=648, llr=bufs["llr"], want_aux=False)
        ms_alt = ta.ms() / 5
        other = {"precision": alt, "kernel": sim.ofdm.last_kernel, "ms_per_launch": ms_alt,
                 "achieved": (4 * L + 4 * 648) * B / (ms_alt * 1e-3) / 1e9,
                 "fer": [round(float(r[1]) / max(int(r[0]), 1), 5) for r in c_alt.cpu().numpy()],
                 "frame_error_count_difference": [int(a[1]) * 1 - int(round(b[1] / args.steps)) for a, b in zip(c_alt.cpu().numpy(), c)]
                 if world == 1 else None}
        sim.ofdm.set_precision(args.precision)

    # ---- the Monte-Carlo loop itself: channel generation INSIDE the timed region (LinkSim.sweep: descriptors -> channel kernel ->
    #      demod -> LDPC -> count per batch), on this workload's channel and, for the headline, on Watterson 'good' as well
    sweeps = {}
    if args.sweep_batches > 0:
        import ctypes as C
        wf = {"ofdm": capi.WF_OFDM, "dpsk": capi.WF_DPSK}[wl["kind"]]
        for chname in ([wl["channel"]] + (["good"] if args.workload == "m1" else [])):
            # the C++ driver (pu_linksim_run, csrc/sweep.cu): descriptors -> channel kernel -> demod -> LDPC -> count per batch, two
            # batches in flight; this rank's share of the (SNR point, seed block) units, no exchange but the final counter sum
            mode = capi.sweep_mode(wf, cfg, rate, pb, chname, snr_points[0], snr_points[1] - snr_points[0], n_snr,
                                   peak=0.5 if wl["kind"] == "dpsk" else 0.0, precision=args.precision if wl["kind"] == "ofdm" else "exact")
            sw = capi.Sweep([mode], trials_per_point=fpp * args.sweep_batches * world, block_trials=min(4096, fpp), pool=POOL if wl["kind"] == "ofdm" else 16,
                            rank=rank, world=world, batch_bytes=B * L * 4)
            capi.Sweep([mode], trials_per_point=fpp, block_trials=min(4096, fpp), pool=POOL if wl["kind"] == "ofdm" else 16).run(ctx)   # warm-up
            barrier()
            cs_h, stt = sw.run(ctx)
            dt = max_over_ranks(stt.seconds)
            cs = torch.from_numpy(cs_h.astype(np.int64)).to(dev)
            linksim.allreduce_counters(cs)
            fr = torch.tensor([float(stt.frames_run)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(fr)
            # the channel kernel alone on one batch of the same shape
            chcfg = linksim.ChannelConfig()
            capi.check(capi.lib().pu_channel_preset(capi.CHANNELS[chname], C.byref(chcfg)))
            rxb = torch.empty_like(rx)
            linksim.channel_apply(ctx, chcfg, sim.tx_pool, batch["tx_index"], batch["noise_std"], batch["seed"], rxb)
            with Timer(torch) as tc:
                linksim.channel_apply(ctx, chcfg, sim.tx_pool, batch["tx_index"], batch["noise_std"], batch["seed"], rxb)
            ms_ch = tc.ms()
            del rxb
            kern = "awgn_kernel" if chname == "awgn" else "channel_kernel"
            sweeps[chname] = {"value": float(fr.item()) / dt, "value_excluding_setup": float(fr.item()) / max(dt - stt.setup_seconds, 1e-9), "unit": UNIT, "api": "pu_linksim_run (C++ driver, channel generation inside)",
                              "frames": int(fr.item()), "seconds": dt, "setup_seconds": stt.setup_seconds, "gpu_wait_seconds": stt.wait_seconds, "host_fill_seconds": stt.fill_seconds, "units_per_rank": int(stt.units_run),
                              "channel_kernel": {"kernel": kern, "ms_per_launch": ms_ch, "bound": "hbm",
                                                 "algorithmic_bytes_per_frame": 4 * L,
                                                 "achieved": 4 * L * B / (ms_ch * 1e-3) / 1e9},
                              "fer": [round(float(r[1]) / max(int(r[0]), 1), 5) for r in cs.cpu().numpy()]}
            barrier()

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D of the samples, D2H of the decoded bytes/flags
    e2e = None
    if wl["kind"] == "ofdm":
        rx_host = torch.empty((B, L), dtype=torch.float32, pin_memory=True)
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
        dt = max_over_ranks(time.perf_counter() - t0)
        tb1 = ctx.transfer_bytes    # bytes the library actually moved (only the FFT windows of the symbols the kernel reads cross PCIe)
        step()                      # device-path outputs of the same mode for the comparison below
        torch.cuda.synchronize()
        e2e_matches = bool((info_h == bufs["info"].cpu().numpy()).all() and (ok_h == bufs["ok"].cpu().numpy()).all())
        h2d = (tb1[0] - tb0[0]) // e2e_steps
        e2e = {"value": world * B * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": world * h2d,
               "d2h_bytes_per_step": world * (tb1[1] - tb0[1]) // e2e_steps, "steps": e2e_steps,
               "host_buffer_bytes_per_step": world * B * L * 4, "api": "pu_receive_decode_batch(PU_MEM_HOST)",
               "matches_device_path": e2e_matches, "h2d_gbs_per_gpu": h2d * e2e_steps / dt / 1e9}
    barrier()
    clocks = sampler.stop(world) if sampler else None

    if rank == 0:
        peak, peak_src = measured_peak()
        alg = 4 * L + 4 * 648
        ach = alg * B / (ms_demod * 1e-3) / 1e9
        traffic, traffic_src = traffic_of(demod_kernel)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(base_config(fpp, world) if args.workload == "m1" else
                               {"workload": "%s x %d frames/point/GPU, demod+demap+LDPC(flooding min-sum, <=50 it)+error count" % (wl["name"], fpp),
                                "frames_per_step_per_gpu": B, "snr_points_db": [snr_points[0], snr_points[-1], snr_points[1] - snr_points[0]],
                                "channel": wl["channel"], "payload_bytes": pb,
                                "parallelism": "frames sharded over %d GPU(s), one counter all-reduce" % world},
                               l2="inputs larger than L2: %.0f MB of samples per step per GPU, no flush" % (B * L * 4 / 1e6),
                               precision=("%s (pu_ofdm_set_precision; see `other_precision` for the same inputs through the other arithmetic)"
                                          % sim.ofdm.precision) if wl["kind"] == "ofdm" else "exact"),
                "info_bits_per_s": value * wl["info_bits"], "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"kernel": demod_kernel, "bound": "hbm", "achieved": ach, "peak": peak,
                             "unit": "GB/s", "frac": ach / peak, "traffic": traffic * B if traffic else None,
                             "traffic_source": traffic_src, "peak_source": peak_src,
                             "algorithmic_bytes_per_frame": alg, "ms_per_launch": ms_demod},
                "stages_ms": {demod_kernel: ms_demod, "ldpc_flood_kernel": ms_ldpc, "count_errors_kernel": ms_count},
                "ldpc": {"codewords_per_s": B / (ms_ldpc * 1e-3), "avg_iterations_run": iters_run,
                         "edge_updates_per_s": 2 * sim.ldpc.num_edges * iters_run * B / (ms_ldpc * 1e-3),
                         "hbm_gbs": (4 * 648 + sim.ldpc.info_bytes + 5) * B / (ms_ldpc * 1e-3) / 1e9},
                "fer": [round(float(r[1]) / max(int(r[0]), 1), 5) for r in c],
                "frames_counted": int(c[:, 0].sum())}
        if e2e:
            line["e2e"] = e2e
        if sustained:
            w0, w1 = sustained.pop("window")
            sustained["clocks"] = sampler.window(w0, w1, world) if sampler else None
            line["sustained"] = sustained
        if other:
            other["frac"] = other["achieved"] / peak
            line["other_precision"] = other
        if sweeps:
            for v in sweeps.values():
                v["channel_kernel"]["frac"] = v["channel_kernel"]["achieved"] / peak
            line["sweep"] = sweeps
        if world == 1 and not args.no_cpu_baseline and args.workload == "m1":
            os.sched_setaffinity(0, cpus_before)      # the reference arm uses every host core
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_ldpc(args, ctx, dev, world, rank):
    """BASELINE.json config 2: the LDPC decoder alone on a batch of 1M codewords at R1/4, R1/2, R3/4, R5/6 (BPSK over AWGN at the
    rate's waterfall: about half of the codewords converge).  A step decodes the four batches once.  The decoder is bound by SM
    issue / ALU and the shared-memory pipe, not by HBM (north star): `roofline` is the HBM view the contract asks for (tiny by
    design), `sm` carries edge-message updates/s -- the unit the kernel is optimised in."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from projectultra_b200 import capi
    B = args.ldpc_codewords
    rates = [("R1_4", 1.12), ("R1_2", 0.71), ("R3_4", 0.57), ("R5_6", 0.58)]
    decs, llrs, outs = [], [], []
    for name, sigma in rates:
        r = getattr(capi, name)
        dec = capi.LdpcDecoder(ctx, r)
        rng = np.random.default_rng(100 + r + 1000 * rank)
        cws = np.stack([np.unpackbits(capi.ldpc_encode(r, rng.integers(0, 256, dec.info_bytes, dtype=np.uint8)))[:648] for _ in range(64)])
        bits = torch.from_numpy(cws.astype(np.float32)).to(dev)
        g = torch.Generator(device=dev)
        g.manual_seed(1 + rank)
        llr = torch.empty((B, 648), dtype=torch.float32, device=dev)
        for off in range(0, B, 1 << 18):                 # bounded temporaries
            n = min(1 << 18, B - off)
            y = (1 - 2 * bits)[(torch.arange(n, device=dev) + off) % 64] + sigma * torch.randn((n, 648), device=dev, generator=g)
            llr[off:off + n] = torch.clamp(2 * y / sigma ** 2, -10, 10)
        decs.append(dec); llrs.append(llr)
        outs.append((torch.empty((B, dec.info_bytes), dtype=torch.uint8, device=dev), torch.empty(B, dtype=torch.uint8, device=dev),
                     torch.empty(B, dtype=torch.int32, device=dev)))

    def step(ev=None):
        for i, (dec, llr, o) in enumerate(zip(decs, llrs, outs)):
            if ev is not None:
                ev[i].record()
            dec.decode_batch(llr, *o)
        if ev is not None:
            ev[len(decs)].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(os.path.join(ROOT, "gpurun_out", "bench_clocks.csv")) if rank == 0 else None
    if sampler:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(decs) + 1)] for _ in range(args.steps)]
    n0 = ctx.kernel_launches
    barrier()
    with Timer(torch) as tm:
        for k in range(args.steps):
            step(evs[k])
    barrier()
    t = torch.tensor([tm.ms()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    clocks = sampler.stop(world) if sampler else None
    if rank == 0:
        peak, peak_src = measured_peak()
        per_rate, upd_total = {}, 0.0
        for i, ((name, sigma), dec, o) in enumerate(zip(rates, decs, outs)):
            ms = sum(e[i].elapsed_time(e[i + 1]) for e in evs) / args.steps
            it_run = float((o[2].float() + o[1].float()).clamp(max=50).mean().item())
            upd = 2 * dec.num_edges * it_run * B
            upd_total += upd
            per_rate[name] = {"sigma": sigma, "ms_per_launch": ms, "codewords_per_s": B / (ms * 1e-3), "converged": float(o[1].float().mean().item()),
                              "avg_iterations_run": it_run, "edge_updates_per_s": upd / (ms * 1e-3),
                              "hbm_gbs": (4 * 648 + dec.info_bytes + 5) * B / (ms * 1e-3) / 1e9}
        value = world * 4 * B * args.steps / (ms_total * 1e-3)
        alg = sum((4 * 648 + d.info_bytes + 5) for d in decs) * B
        ach = alg / (ms_total / args.steps * 1e-3) / 1e9
        traffic, traffic_src = traffic_of("ldpc_flood_reg_kernel")
        line = {"metric": "decoded codewords/sec (LDPC alone)", "value": value, "unit": "codewords/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "config 2: LDPC flooding min-sum decoder alone, %d codewords per rate per GPU at R1/4, R1/2, R3/4, R5/6, "
                                       "BPSK/AWGN at each rate's waterfall, <= 50 iterations, syndrome stop" % B,
                           "codewords_per_step_per_gpu": 4 * B, "l2": "inputs larger than L2: %.1f GB of LLRs per step, no flush" % (4 * B * 2592 / 1e9)},
                "gpu_launches": int(ctx.kernel_launches - n0), "clocks": clocks,
                "roofline": {"kernel": "ldpc_flood_reg_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": traffic * 4 * B if traffic else None, "traffic_source": traffic_src, "peak_source": peak_src,
                             "note": "not the binding resource: the decoder is SM-issue / shared-memory-pipe bound (see `sm`)"},
                "sm": {"edge_updates_per_s": upd_total / (ms_total / args.steps * 1e-3), "per_rate": per_rate,
                       "ncu": "profiles/: smsp__issue_active, sm__inst_executed_pipe_alu, l1tex__data_pipe_lsu_wavefronts_mem_shared of the same kernel"}}
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
    ap.add_argument("--workload", default="m1", choices=["m1", "ldpc", "m3", "dpsk"],
                    help="m1 = the headline (BASELINE.json configs[1]'s mode); ldpc / m3 / dpsk = configs 2 / 3 / 4")
    ap.add_argument("--precision", default="fast", choices=["fast", "exact"],
                    help="arithmetic of the OFDM kernels that have an FMA form (pu_ofdm_set_precision); the other one is timed on the "
                         "same inputs and reported under other_precision")
    ap.add_argument("--sustain-seconds", type=float, default=2.5, help="length of the back-to-back `sustained` leg (0 = skip)")
    ap.add_argument("--sweep-batches", type=int, default=200, help="batches of the `sweep` leg (channel inside the timed region; 0 = skip)")
    ap.add_argument("--ldpc-codewords", type=int, default=1 << 20, help="--workload ldpc: codewords per rate per GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        if args.workload != "m1":
            if int(os.environ.get("RANK", "0")) == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the reference arm is implemented for --workload m1 (the headline) only"}))
            return 0
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
