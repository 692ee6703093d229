"""N>1 host logic on CPU: two gloo ranks take disjoint trial shards, fill counter tables and all-reduce them; the
result equals the single-process table (the path's only exchange step)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["PU_ROOT"])
    from projectultra_b200 import linksim
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["PU_PORT"],
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    n_snr, trials = 5, 1001
    mine = linksim.shard_trials(trials, rank, world)
    c = torch.zeros((n_snr, 6), dtype=torch.int64)
    for s in range(n_snr):                       # a deterministic stand-in for the GPU counting kernel
        h = (mine * 2654435761 + s * 40503) % 97
        c[s, 0] = len(mine); c[s, 1] = int((h < 10 * (n_snr - s)).sum()); c[s, 2] = int(h.sum()); c[s, 3] = 320 * len(mine)
        c[s, 5] = int((h % 50).sum())
    linksim.allreduce_counters(c)
    if rank == 0:
        print("RESULT " + json.dumps(c.tolist()))
    dist.destroy_process_group()
""")


def run(world, port):
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), PU_PORT=str(port), PU_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0]
    import json
    return np.array(json.loads(line[7:]))


def test_two_rank_gloo_allreduce_equals_single_process():
    one = run(1, 29611)
    two = run(2, 29612)
    assert (one == two).all()
    assert one[:, 0].tolist() == [1001] * 5


def test_shards_are_disjoint_and_exhaustive():
    sys.path.insert(0, ROOT)
    from projectultra_b200 import linksim
    for world in (1, 2, 3, 8):
        allt = np.concatenate([linksim.shard_trials(1000, r, world) for r in range(world)])
        assert sorted(allt.tolist()) == list(range(1000))
    lo, hi = linksim.wilson_interval(5, 100)
    assert lo < 0.05 < hi and linksim.wilson_interval(0, 0) == (0.0, 1.0)
