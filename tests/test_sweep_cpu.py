"""Host logic of the sweep driver (csrc/sweep.cu: pu_sweep_*, BASELINE.json config 5) without a GPU: unit enumeration, the
cost-weighted partitioner (disjoint, exhaustive, deterministic, balanced, resume-aware), the payload generator, Wilson intervals,
channel presets; and a world-size-2 gloo run in which every rank computes its own share of the partition and the per-unit
counters are all-reduced to the single-process table."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def table(capi):
    m1 = capi.ModemConfig(48000, 1500, 512, 30, 1, 4, 2, 0, capi.DQPSK, capi.R1_2, 40.0, 0.0)
    m3 = capi.ModemConfig(48000, 1500, 1024, 59, 1, 0, 4, 1, capi.QAM32, capi.R3_4, 40.0, 0.0)
    return [capi.sweep_mode(capi.WF_OFDM, m1, capi.R1_2, 40, "awgn", -4, 1, 13, precision="fast"),
            capi.sweep_mode(capi.WF_OFDM, m3, capi.R3_4, 60, "good", 8, 1, 13),
            capi.sweep_mode(capi.WF_OFDM_SC, m1, capi.R1_2, 40, "awgn", 0, 2, 6, peak=0.5),
            capi.sweep_mode(capi.WF_DPSK_ACQ, capi.dpsk_config(1, 384), capi.R1_4, 20, "poor", -11, 2, 15, peak=0.5),
            capi.sweep_mode(capi.WF_MCDPSK, capi.mcdpsk_config(8, 2), capi.R1_2, 40, "moderate", -2, 2, 8)]


def test_units_enumerate_the_whole_grid():
    from projectultra_b200 import capi
    sw = capi.Sweep(table(capi), trials_per_point=10000, block_trials=4096)
    assert sw.n_points == 13 + 13 + 6 + 15 + 8 and sw.n_units == sw.n_points * 3
    seen = {}
    for u in range(sw.n_units):
        m, s, t0, nt = sw.unit(u)
        assert nt == (4096 if t0 < 8192 else 10000 - 8192)
        seen.setdefault((m, s), []).append((t0, nt))
    assert len(seen) == sw.n_points
    for (m, s), blocks in seen.items():
        assert s < sw.modes[m].n_snr and sum(nt for _, nt in blocks) == 10000
        assert sorted(t0 for t0, _ in blocks) == [0, 4096, 8192]


def test_partition_is_disjoint_exhaustive_deterministic_and_balanced():
    from projectultra_b200 import capi
    for world in (1, 2, 3, 8):
        sw = capi.Sweep(table(capi), trials_per_point=100000, world=world)
        owner, cost = sw.partition()
        assert (owner < world).all() and (cost > 0).all()
        owner2, _ = capi.Sweep(table(capi), trials_per_point=100000, world=world, rank=world - 1).partition()
        assert (owner == owner2).all()                       # every rank computes the same assignment
        load = np.bincount(owner, weights=cost, minlength=world)
        assert load.max() / load.mean() < 1.02, load / load.mean()
    # low-SNR units are estimated dearer than high-SNR ones of the same mode; acquisition dearer than genie timing
    sw = capi.Sweep(table(capi), trials_per_point=4096, block_trials=4096)
    _, cost = sw.partition()
    assert cost[0] > cost[12] and cost[26] > 20 * cost[0]


def test_partition_skips_finished_units_and_rebalances():
    from projectultra_b200 import capi
    sw = capi.Sweep(table(capi), trials_per_point=20000, world=4)
    rng = np.random.default_rng(3)
    done = (rng.random(sw.n_units) < 0.4).astype(np.uint8)
    owner, cost = sw.partition(done)
    assert (owner[done == 1] == 0xFFFFFFFF).all() and (owner[done == 0] < 4).all()
    load = np.bincount(owner[done == 0], weights=cost[done == 0], minlength=4)
    assert load.max() / load.mean() < 1.05


def test_payloads_wilson_and_presets():
    from projectultra_b200 import capi, linksim
    sw = capi.Sweep(table(capi), trials_per_point=100)
    a, b = sw.payload(0, 5), sw.payload(0, 5)
    assert (a == b).all() and len(a) == 40 and not (a == sw.payload(0, 6)).all() and not (a == sw.payload(1, 5)[:40]).all()
    for e, n in ((0, 0), (0, 100), (5, 100), (100, 100), (37, 12345)):
        assert np.allclose(capi.wilson_interval(e, n), linksim.wilson_interval(e, n), atol=1e-12)
    import ctypes as C
    for name, idx in capi.CHANNELS.items():
        got = linksim.ChannelConfig()
        capi.check(capi.lib().pu_channel_preset(idx, C.byref(got)))
        want = linksim.channel_preset(name.replace("itu_", ""))       # itu_r_f1487:: presets carry the ccir:: numbers (hf_channel.hpp:402-487)
        assert bytes(got) == bytes(want), name


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["PU_ROOT"]); sys.path.insert(0, os.path.join(os.environ["PU_ROOT"], "tests"))
    from projectultra_b200 import capi
    from test_sweep_cpu import table
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["PU_PORT"],
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    sw = capi.Sweep(table(capi), trials_per_point=9000, block_trials=2048, rank=rank, world=world)
    owner, cost = sw.partition()
    mine = np.flatnonzero(owner == rank)
    c = torch.zeros((sw.n_points, 6), dtype=torch.int64)
    point0 = np.cumsum([0] + [m.n_snr for m in sw.modes])
    for u in mine:                                  # a deterministic stand-in for the kernels: counters are functions of the unit
        m, s, t0, nt = sw.unit(int(u))
        t = np.arange(t0, t0 + nt)
        h = (t * 2654435761 + s * 40503 + m * 977) % 97
        row = point0[m] + s
        c[row, 0] += nt; c[row, 1] += int((h < 40 - 2 * s).sum()); c[row, 2] += int(h.sum()); c[row, 3] += 320 * nt; c[row, 5] += int((h % 50).sum())
    owned = torch.zeros(sw.n_units, dtype=torch.int64); owned[mine] = 1
    dist.all_reduce(c); dist.all_reduce(owned)
    if rank == 0:
        print("RESULT " + json.dumps({"c": c.tolist(), "owned_once": bool((owned == 1).all()), "units": int(sw.n_units)}))
    dist.destroy_process_group()
""")


def run(world, port):
    import json
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), PU_PORT=str(port), PU_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    return json.loads([l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0][7:])


def test_two_gloo_ranks_cover_the_grid_once_and_sum_to_the_single_process_table():
    one, two = run(1, 29631), run(2, 29632)
    assert one["owned_once"] and two["owned_once"]
    assert one["c"] == two["c"]
    assert all(row[0] == 9000 for row in one["c"])


def test_ctypes_mirrors_have_the_c_structs_sizes():
    """The Python mirrors of the public PODs are written by hand; pu_abi_sizes reports what the C side compiled."""
    import ctypes as C
    from projectultra_b200 import capi, linksim
    out = (C.c_uint32 * 8)()
    assert capi.lib().pu_abi_sizes(out) == 7
    want = [capi.ModemConfig, capi.DpskConfig, capi.McDpskConfig, linksim.ChannelConfig, capi.SweepMode, capi.SweepDesc, capi.SweepStats]
    assert [int(v) for v in out[:7]] == [C.sizeof(t) for t in want]
