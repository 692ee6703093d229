#!/bin/bash
OUT=gpurun_out/r10; mkdir -p $OUT
( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; tail -25 $OUT/pytest_gpu.log | cut -c1-250
for m in m1 m3 m1qam16; do python tools/ofdm_quick_bench.py 4096 $m > $OUT/quick_$m.log 2>&1; tail -1 $OUT/quick_$m.log; done
PU_OFDM_NO_WARPG=1 python tools/ofdm_quick_bench.py 4096 m3 > $OUT/quick_m3_cta.log 2>&1; tail -1 $OUT/quick_m3_cta.log
