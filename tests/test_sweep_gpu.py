"""GPU tests of the C++ sweep driver (pu_linksim_run, csrc/sweep.cu) against the Python link simulator that the parity tests of
test_linksim_gpu.py pin frame by frame to the oracle: same payload pool, same seeds -> identical counter tables, for one mode of
every waveform family; two ranks emulated on one GPU sum to the single-rank table; an interrupted sweep resumes from its manifest
to the uninterrupted totals."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from projectultra_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def small_table(capi):
    m1 = capi.ModemConfig(48000, 1500, 512, 30, 1, 4, 2, 0, capi.DQPSK, capi.R1_2, 40.0, 0.0)
    m1q = capi.ModemConfig(48000, 1500, 512, 30, 1, 4, 2, 1, capi.QAM16, capi.R1_2, 40.0, 0.0)
    return [capi.sweep_mode(capi.WF_OFDM, m1, capi.R1_2, 40, "awgn", -3, 1.5, 5, precision="fast"),
            capi.sweep_mode(capi.WF_OFDM, m1q, capi.R1_2, 40, "good", 8, 3, 4),
            capi.sweep_mode(capi.WF_DPSK, capi.dpsk_config(1, 384), capi.R1_4, 20, "poor", -6, 4, 4, peak=0.5),
            capi.sweep_mode(capi.WF_MCDPSK, capi.mcdpsk_config(8, 2), capi.R1_2, 40, "moderate", 0, 3, 4)]


def python_counters(ctx, sw, mi, trials):
    """The same (mode, SNR point, trial) grid through projectultra_b200.linksim.LinkSim."""
    import torch
    from projectultra_b200 import capi, linksim
    m = sw.modes[mi]
    pool = sw.desc.pool
    payloads = np.stack([sw.payload(mi, i) for i in range(pool)])
    cfg = m.ofdm if m.waveform <= capi.WF_OFDM_CHIRP else m.dpsk if m.waveform <= capi.WF_DPSK_ACQ else m.mcdpsk
    chan = [k for k, v in capi.CHANNELS.items() if v == m.channel][0]
    sim = linksim.LinkSim(ctx, cfg, chan, payload_bytes=m.payload_bytes, pool=pool, code_rate=m.code_rate, peak=m.peak or None,
                          precision="fast" if m.precision else "exact", payloads=payloads)
    snrs = m.snr_points
    si = np.repeat(np.arange(len(snrs), dtype=np.int64), trials)
    tr = np.tile(np.arange(trials, dtype=np.int64), len(snrs))
    batch = sim.make_batch(snrs, si, tr, base_seed=0xB200 | (mi << 16))      # (mode << 56) of the driver's seed rule
    c = torch.zeros((len(snrs), 6), dtype=torch.int64, device="cuda")
    sim.run_batch(batch, c)
    torch.cuda.synchronize()
    return c.cpu().numpy().astype(np.uint64)


def test_driver_matches_python_linksim_mode_by_mode(ctx):
    from projectultra_b200 import capi
    trials = 96
    sw = capi.Sweep(small_table(capi), trials_per_point=trials, block_trials=40, pool=8)
    counters, st = sw.run(ctx)
    assert st.units_run == sw.n_units and st.frames_run == trials * sw.n_points and st.units_resumed == 0
    at = 0
    for mi, m in enumerate(sw.modes):
        want = python_counters(ctx, sw, mi, trials)
        got = counters[at:at + m.n_snr]
        assert (got == want).all(), (mi, got.tolist(), want.tolist())
        assert (got[:, 0] == trials).all()
        at += m.n_snr
    # a waterfall: the first SNR point of the fast OFDM mode loses frames, the last one does not
    assert counters[0, 1] > counters[4, 1] == 0


def test_two_ranks_on_one_gpu_sum_to_the_single_rank_table(ctx):
    from projectultra_b200 import capi
    one, _ = capi.Sweep(small_table(capi), trials_per_point=64, block_trials=16, pool=8).run(ctx)
    parts = []
    for r in range(2):
        c, st = capi.Sweep(small_table(capi), trials_per_point=64, block_trials=16, pool=8, rank=r, world=2).run(ctx)
        assert 0 < st.units_run < st.units_total and abs(st.busy_cost / st.total_cost - 0.5) < 0.05
        parts.append(c)
    assert (parts[0] + parts[1] == one).all()


def test_interrupted_sweep_resumes_from_its_manifest(ctx, tmp_path):
    from projectultra_b200 import capi
    d = str(tmp_path / "manifest")
    full, _ = capi.Sweep(small_table(capi), trials_per_point=48, block_trials=16, pool=8).run(ctx)
    c1, s1 = capi.Sweep(small_table(capi), trials_per_point=48, block_trials=16, pool=8, manifest_dir=d, max_units=7, run_id=1).run(ctx)
    assert s1.units_run == 7 and s1.units_resumed == 0
    shard = [f for f in os.listdir(d) if f.startswith("shard-")]
    assert len(shard) == 1
    with open(os.path.join(d, shard[0]), "a") as f:
        f.write("12 16 3 ")                                                  # a torn record (killed while appending): ignored
    # resume with a different world size: finished units are skipped whoever ran them, rank 0 returns their counters.  Both ranks of
    # the new launch (run_id 2) must see the same finished set although rank 1 starts after rank 0 has appended its units.
    tot = np.zeros_like(full)
    for r in range(2):
        c, st = capi.Sweep(small_table(capi), trials_per_point=48, block_trials=16, pool=8, manifest_dir=d, rank=r, world=2, run_id=2).run(ctx)
        assert st.units_resumed == 7 and st.units_run > 0
        tot += c
    assert (tot == full).all()
    # everything is in the manifest now: nothing left to run, the table comes back from the files
    c3, s3 = capi.Sweep(small_table(capi), trials_per_point=48, block_trials=16, pool=8, manifest_dir=d, run_id=3).run(ctx)
    assert s3.units_run == 0 and s3.units_resumed == s3.units_total and (c3 == full).all()
    # a different table must not resume from this directory
    with pytest.raises(capi.PuError):
        capi.Sweep(small_table(capi), trials_per_point=49, block_trials=16, pool=8, manifest_dir=d, run_id=4).run(ctx)


def test_acquired_modes_need_the_tools_silence_behind_the_frame(ctx):
    """tools/test_iwaveform.cpp:396-459 surrounds every frame with silence.  Without a tail, the delayed path of a Watterson channel pushes
    the end of the last symbol out of the buffer: the receiver then reports fewer than 648 soft bits and the frame counts as lost at any
    SNR (the 0.99 floor of the first config-5 table, profiles/r2_v44_config5_sweep.md).  pu_sweep_mode.lead_samples / tail_samples."""
    from projectultra_b200 import capi
    mc = capi.mcdpsk_config(8, 2)
    m1 = capi.ModemConfig.from_buffer_copy(bytes(__import__("refapi").config_m1(__import__("refapi").DQPSK, __import__("refapi").R1_4)))
    fer = {}
    for tail in (0, 2400):
        modes = [capi.sweep_mode(capi.WF_MCDPSK_CHIRP, mc, capi.R1_4, 20, "moderate", 20, 5, 2, peak=0.5, lead_samples=480 if tail else 0, tail_samples=tail),
                 capi.sweep_mode(capi.WF_OFDM_CHIRP, m1, capi.R1_4, 20, "good", 20, 5, 2, peak=0.5, precision="fast", lead_samples=480 if tail else 0,
                                 tail_samples=tail)]
        counters, _ = capi.Sweep(modes, trials_per_point=128, block_trials=64, pool=8).run(ctx)
        fer[tail] = counters[:, 1].astype(np.float64) / counters[:, 0]
    print("FER without / with silence:", fer[0].round(3).tolist(), fer[2400].round(3).tolist())
    assert (fer[0] > 0.7).all()                       # every frame short of a codeword
    assert (fer[2400] < 0.45).all() and fer[2400][2:].max() < 0.1      # MC-DPSK: the moderate channel's fades; OFDM_CHIRP R1/4 on good: clean


def test_tuning_error_rows_of_the_regression_matrix(ctx):
    """tests/regression_matrix.sh:139-243 runs the chirp-acquired waveforms with --cfo 0 / 30 / 50: the tools' FFT-Hilbert injector on the
    clean TX audio (pu_sweep_mode.cfo_hz -> pu_tools_apply_cfo), estimated by the dual chirp and removed by the receiver.  The rows the
    matrix expects to pass ("--snr 17 --cfo 30/50 awgn ofdm_chirp", "--snr 5 --cfo 30 awgn mc_dpsk") decode here too, and a tuning error
    of 30 Hz without the chirp's estimate would not (the same frames through the genie-timed receiver lose every frame)."""
    import refapi as R
    from projectultra_b200 import capi
    m1 = capi.ModemConfig.from_buffer_copy(bytes(R.config_m1(R.DQPSK, R.R1_2)))
    mc = capi.mcdpsk_config(8, 2)
    kw = dict(peak=0.5, lead_samples=480, tail_samples=2400)
    modes = [capi.sweep_mode(capi.WF_OFDM_CHIRP, m1, capi.R1_2, 40, "awgn", 17, 1, 1, precision="fast", cfo_hz=c, **kw) for c in (0.0, 30.0, 50.0)]
    modes += [capi.sweep_mode(capi.WF_MCDPSK_CHIRP, mc, capi.R1_2, 40, "awgn", 5, 1, 1, cfo_hz=c, **kw) for c in (0.0, 30.0, -30.0)]
    modes += [capi.sweep_mode(capi.WF_MCDPSK, mc, capi.R1_2, 40, "awgn", 5, 1, 1, cfo_hz=c, peak=0.5) for c in (0.0, 30.0)]      # no chirp: no estimate
    counters, _ = capi.Sweep(modes, trials_per_point=128, block_trials=64, pool=8).run(ctx)
    fer = counters[:, 1].astype(np.float64) / counters[:, 0]
    print("FER  OFDM_CHIRP 17 dB cfo 0/30/50:", fer[:3].round(3).tolist(), " MC-DPSK behind the chirp 5 dB cfo 0/30/-30:", fer[3:6].round(3).tolist(),
          " MC-DPSK genie-timed cfo 0/30:", fer[6:].round(3).tolist())
    assert (fer[:6] < 0.1).all()
    assert fer[6] < 0.1 and fer[7] > 0.9


def test_fresh_payload_per_trial_matches_the_pool_statistically(ctx):
    """pu_sweep_mode.fresh_payloads: payload -> LDPC encode -> modulate on the GPU for every trial of a batch (tools/test_dpsk_snr.cpp:40-60,
    tools/test_mode_snr.cpp:40-60 draw a new payload per trial) instead of the pool of host-built TX waveforms.  Same FER within the
    binomial spread, a waterfall in every mode, reproducible, and two ranks sum to the single-rank table."""
    import refapi as R
    from projectultra_b200 import capi
    m1 = capi.ModemConfig.from_buffer_copy(bytes(R.config_m1(R.DQPSK, R.R1_2)))
    m1q = capi.ModemConfig.from_buffer_copy(bytes(R.config_m1(R.QAM16, R.R1_2)))

    def table(fresh):
        return [capi.sweep_mode(capi.WF_OFDM, m1, capi.R1_2, 40, "awgn", -1, 1.5, 4, precision="fast", fresh_payloads=fresh),
                capi.sweep_mode(capi.WF_OFDM, m1q, capi.R1_2, 40, "good", 10, 4, 3, fresh_payloads=fresh),
                capi.sweep_mode(capi.WF_DPSK, capi.dpsk_config(1, 384), capi.R1_4, 20, "awgn", -22, 2, 4, peak=0.5, fresh_payloads=fresh),
                capi.sweep_mode(capi.WF_DPSK_ACQ, capi.dpsk_config(1, 384), capi.R1_4, 20, "poor", -20, 4, 3, peak=0.5, lead_samples=480, tail_samples=2400,
                                fresh_payloads=fresh),
                capi.sweep_mode(capi.WF_MCDPSK, capi.mcdpsk_config(8, 2), capi.R1_2, 40, "moderate", 0, 3, 4, peak=0.5, fresh_payloads=fresh)]

    trials = 512
    pool, _ = capi.Sweep(table(False), trials_per_point=trials, block_trials=128, pool=32).run(ctx)
    fresh, st = capi.Sweep(table(True), trials_per_point=trials, block_trials=128, pool=32).run(ctx)
    again, _ = capi.Sweep(table(True), trials_per_point=trials, block_trials=128, pool=32).run(ctx)
    assert (fresh == again).all()
    parts = [capi.Sweep(table(True), trials_per_point=trials, block_trials=128, pool=32, rank=r, world=2).run(ctx)[0] for r in range(2)]
    assert (parts[0] + parts[1] == fresh).all()
    assert (fresh[:, 0] == trials).all() and (fresh[:, 3] == pool[:, 3]).all()            # frames, payload bits compared
    fp, ff = pool[:, 1] / trials, fresh[:, 1] / trials
    print("FER pool :", fp.round(3).tolist())
    print("FER fresh:", ff.round(3).tolist())
    sigma = np.sqrt(np.maximum(fp * (1 - fp), 0.02) * 2 / trials)
    assert (np.abs(fp - ff) < 5 * sigma + 0.02).all(), (fp - ff).tolist()
    at = 0
    for m in table(True):
        assert ff[at] > ff[at + m.n_snr - 1] or ff[at] == 0.0                           # a waterfall (or already clean)
        at += m.n_snr
