"""Pinned host -> device copy ceiling of this box at N ranks (one process per GPU, torchrun): the denominator of the end-to-end
numbers of bench.py (`e2e.h2d_gbs_per_gpu`), which are bound by exactly this copy.

    python tools/h2d_bench.py                                              # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29720 tools/h2d_bench.py

Every rank copies a 1.3 GB pinned buffer (the size of one bench step's sample windows) to its GPU 10 times, all ranks at the same
time (barrier before, max over ranks after); rank 0 prints one JSON line with the per-rank and aggregate GB/s for: one contiguous
cudaMemcpyAsync, the same split over two streams, and the strided window copy of pu_receive_decode_batch (slabs of 4 096 frames,
12 x 512 of 7 332 samples per frame) timed through the library itself with the kernels' share removed (device-path time)."""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sys.path.insert(0, ROOT)
import bench as B
B.bind_near_cpus(local)

n_bytes = 53248 * 12 * 512 * 4
host = torch.empty(n_bytes // 4, dtype=torch.float32, pin_memory=True)
host.normal_()
dst = torch.empty_like(host, device=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=10):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    barrier()
    return n_bytes * reps / float(t.item()) / 1e9


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = host.numel() // 2


def one():
    dst.copy_(host, non_blocking=True)


def two():
    with torch.cuda.stream(s1):
        dst[:half].copy_(host[:half], non_blocking=True)
    with torch.cuda.stream(s2):
        dst[half:].copy_(host[half:], non_blocking=True)


res = {"n_gpus": world, "bytes_per_copy": n_bytes, "contiguous_1_stream_gbs_per_gpu": timed(one), "contiguous_2_streams_gbs_per_gpu": timed(two)}
res["aggregate_gbs"] = world * max(res["contiguous_1_stream_gbs_per_gpu"], res["contiguous_2_streams_gbs_per_gpu"])
if rank == 0:
    try:
        import subprocess
        res["topology"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[-1500:]
    except Exception:   # noqa: BLE001
        pass
    res["cpus_this_rank"] = len(os.sched_getaffinity(0))
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()
