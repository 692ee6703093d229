"""A reduced pass over EVERY kernel family, sized for compute-sanitizer (memcheck / racecheck run 10-100x slower): tools/visit.sh
`sanitize` runs this file under `compute-sanitizer --tool memcheck|racecheck --error-exitcode 9` and the logs go to profiles/.
Without the sanitizer it is a quick smoke test of the same launches (results are checked by the parity tests proper)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from projectultra_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def test_every_kernel_family_launches_cleanly(ctx):
    import torch
    from projectultra_b200 import capi, linksim
    m1 = capi.ModemConfig(48000, 1500, 512, 30, 1, 4, 2, 0, capi.DQPSK, capi.R1_2, 40.0, 0.0)
    m1q = capi.ModemConfig(48000, 1500, 512, 30, 1, 4, 2, 1, capi.QAM16, capi.R1_2, 40.0, 0.0)
    m3 = capi.ModemConfig(48000, 1500, 1024, 59, 1, 0, 4, 1, capi.QAM32, capi.R3_4, 40.0, 0.0)
    m3d = capi.ModemConfig(48000, 1500, 1024, 59, 1, 0, 2, 0, capi.DQPSK, capi.R3_4, 40.0, 0.0)
    ran = []
    # OFDM presynced: exact + fast 512 kernels (TMA ring, per-warp queue), 1024 warp-FFT kernel, general warp-granular kernel with
    # pilots, CFO rotator (speculate-and-verify phase walk), fused deinterleave, odd batch sizes
    for cfg, rate, nb, chan, precs in ((m1, capi.R1_2, 40, "awgn", ("exact", "fast")), (m1q, capi.R1_2, 40, "good", ("exact",)),
                                       (m3, capi.R3_4, 60, "good", ("exact",)), (m3d, capi.R3_4, 60, "awgn", ("exact",))):
        sim = linksim.LinkSim(ctx, cfg, chan, payload_bytes=nb, pool=4, code_rate=rate)
        snrs = [0.0, 6.0, 14.0]
        si = np.repeat(np.arange(3), 7)
        tr = np.tile(np.arange(7), 3)
        batch = sim.make_batch(snrs, si, tr)
        for prec in precs:
            sim.ofdm.set_precision(prec)
            c = torch.zeros((3, 6), dtype=torch.int64, device="cuda")
            rx, info, ok, it = sim.run_batch(batch, c, keep=True)
            ran.append(sim.ofdm.last_kernel)
            sim.ofdm.set_deinterleave(sim.ofdm.bits_per_symbol, 648)
            sim.ofdm.presynced_batch(rx[:5], 2, llr_stride=648)
            sim.ofdm.set_deinterleave(0)
        cfo = torch.full((21,), 3.5, dtype=torch.float32, device="cuda")
        ph = torch.full((21,), -0.7, dtype=torch.float32, device="cuda")
        sim.ofdm.presynced_batch(rx, 2, cfo, ph)
        ran.append(sim.ofdm.last_kernel)
        sim.ofdm.training_cfo_batch(rx)
        # transmitter + host-buffer pipeline
        pay = torch.randint(0, 256, (9, nb), dtype=torch.uint8, device="cuda")
        sim.ofdm.tx_batch(sim.ldpc, pay)
        linksim.receive_decode(sim.ofdm, sim.ldpc, rx.cpu().numpy())
    # LDPC: every rate through the register kernel, protocol frames
    for rate in (capi.R1_4, capi.R1_2, capi.R2_3, capi.R3_4, capi.R5_6):
        dec = capi.LdpcDecoder(ctx, rate)
        llr = torch.randn((33, 648), device="cuda") * 3
        dec.decode_batch(llr)
        dec.frame_decode_batch(torch.randn((4, 3 * 648), device="cuda") * 4, 3)
    # acquisition: Schmidl-Cox, dual chirp (OFDM and MC-DPSK), Barker
    sim = linksim.LinkSim(ctx, m1, "awgn", payload_bytes=40, pool=2, layout="sc", peak=0.5)
    b = sim.make_batch([8.0], np.zeros(3, np.int64), np.arange(3))
    sim.run_batch(b, torch.zeros((1, 6), dtype=torch.int64, device="cuda"))
    sim = linksim.LinkSim(ctx, m1, "awgn", payload_bytes=40, pool=2, layout="chirp", peak=0.5)
    b = sim.make_batch([10.0], np.zeros(2, np.int64), np.arange(2))
    sim.run_batch(b, torch.zeros((1, 6), dtype=torch.int64, device="cuda"))
    sim = linksim.LinkSim(ctx, capi.mcdpsk_config(8, 2), "good", payload_bytes=40, pool=2, layout="chirp", peak=0.5)
    b = sim.make_batch([10.0], np.zeros(2, np.int64), np.arange(2))
    sim.run_batch(b, torch.zeros((1, 6), dtype=torch.int64, device="cuda"))
    sim = linksim.LinkSim(ctx, capi.mcdpsk_config(5, 2), "poor", payload_bytes=40, pool=2)
    b = sim.make_batch([6.0], np.zeros(5, np.int64), np.arange(5))
    sim.run_batch(b, torch.zeros((1, 6), dtype=torch.int64, device="cuda"))
    for acquire in (False, True):
        sim = linksim.LinkSim(ctx, capi.dpsk_config(1, 384), "poor", payload_bytes=20, pool=2, code_rate=capi.R1_4, peak=0.5, acquire=acquire)
        b = sim.make_batch([9.0], np.zeros(3, np.int64), np.arange(3))
        sim.run_batch(b, torch.zeros((1, 6), dtype=torch.int64, device="cuda"))
    # the C++ sweep driver (two batches in flight, counters D2H)
    modes = [capi.sweep_mode(capi.WF_OFDM, m1, capi.R1_2, 40, "awgn", 0, 4, 3, precision="fast"),
             capi.sweep_mode(capi.WF_MCDPSK, capi.mcdpsk_config(8, 2), capi.R1_2, 40, "moderate", 2, 4, 2)]
    counters, st = capi.Sweep(modes, trials_per_point=24, block_trials=8, pool=4).run(ctx)
    torch.cuda.synchronize()
    assert st.frames_run == 24 * 5 and (counters[:, 0] == 24).all()
    assert {"ofdm_diff512_kernel", "ofdm_fast512_kernel", "ofdm_diff_kernel", "ofdm_presynced_warp_kernel"} <= set(ran), ran
