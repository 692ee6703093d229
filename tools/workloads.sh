#!/bin/bash
# Timings of the other §8 configurations next to the bench workload (one GPU): config 2 (LDPC alone, 1M codewords per rate
# and operating point), config 3 / pilot modes (general presynced kernel), acquisition (config 1 as literally specified),
# transmitter.  Usage: bash tools/workloads.sh <outdir>
OUT=${1:-gpurun_out/workloads}; mkdir -p $OUT
python tools/ldpc_quick_bench.py 1048576 > $OUT/ldpc_config2.log 2>&1
for m in m1 m3 m1qam16; do python tools/ofdm_quick_bench.py 4096 $m > $OUT/demod_$m.log 2>&1; done
QB_CHANNEL=good python tools/ofdm_quick_bench.py 4096 m3 > $OUT/demod_m3_good.log 2>&1
python tools/acquire_quick_bench.py 8192 > $OUT/acquire.log 2>&1
python tools/dpsk_acquire_quick_bench.py 2048 > $OUT/dpsk_acquire.log 2>&1
python tools/chirp_quick_bench.py 2048 > $OUT/chirp.log 2>&1
python tools/chirp_quick_bench.py 2048 mcdpsk > $OUT/chirp_mcdpsk.log 2>&1
python tools/frame_quick_bench.py 65536 > $OUT/frames.log 2>&1
python - > $OUT/tx.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from projectultra_b200 import capi
ctx = capi.Context(0)
for name, cfg, rate, nb in (("M1 DQPSK R1/2", capi.ModemConfig(48000, 1500, 512, 30, 1, 4, 2, 0, capi.DQPSK, capi.R1_2, 40.0, 0.0), capi.R1_2, 40),
                            ("M3 32QAM R3/4", capi.ModemConfig(48000, 1500, 1024, 59, 1, 0, 4, 1, capi.QAM32, capi.R3_4, 40.0, 0.0), capi.R3_4, 60)):
    dem = capi.OfdmDemodulator(ctx, cfg); enc = capi.LdpcDecoder(ctx, rate)
    B = 53248
    pay = torch.randint(0, 256, (B, nb), dtype=torch.uint8, device="cuda")
    out = dem.tx_batch(enc, pay); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): dem.tx_batch(enc, pay, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("tx %-14s B=%d L=%d ms=%.3f  %.1f Mframes/s  %.1f GB/s written" % (name, B, out.shape[1], ms, B / ms / 1e3, B * out.shape[1] * 4 / ms / 1e6))
PY
{
echo "## Other workloads (tools/workloads.sh, one B200)"
echo; echo "### Config 2: LDPC alone, 1 048 576 codewords (tools/ldpc_quick_bench.py)"; echo '```'; cat $OUT/ldpc_config2.log; echo '```'
echo; echo "### Demodulator kernels, 53 248 frames (tools/ofdm_quick_bench.py: m1 = headline, m3 = config 3, m1qam16 = pilots/2)"; echo '```'
for f in demod_m1 demod_m3 demod_m3_good demod_m1qam16; do tail -1 $OUT/$f.log | sed 's/^ *//'; done; echo '```'
echo; echo "### Acquisition (config 1 as literally specified), 8 192 frames of 10 124 samples (tools/acquire_quick_bench.py)"; echo '```'; cat $OUT/acquire.log; echo '```'
echo; echo "### DPSK Barker acquisition (config 4 as literally specified), 2 048 frames of 139 392 samples (tools/dpsk_acquire_quick_bench.py)"; echo '```'; cat $OUT/dpsk_acquire.log; echo '```'
echo; echo "### Dual-chirp synchronisation + presynced demodulation of OFDM_CHIRP frames (tools/chirp_quick_bench.py)"; echo '```'; cat $OUT/chirp.log; echo '```'
echo; echo "### Dual-chirp synchronisation + Hilbert-FIR correction + processGotChirp of MC-DPSK frames (tools/chirp_quick_bench.py B mcdpsk)"; echo '```'; cat $OUT/chirp_mcdpsk.log; echo '```'
echo; echo "### Protocol-v2 multi-codeword frames: RxPipeline::decodeFrame for 65 536 five-codeword R1/2 frames (tools/frame_quick_bench.py)"; echo '```'; cat $OUT/frames.log; echo '```'
echo; echo "### Transmitter (pu_ofdm_tx_batch)"; echo '```'; cat $OUT/tx.log; echo '```'
} > $OUT/workloads.md
cat $OUT/workloads.md | tail -40
