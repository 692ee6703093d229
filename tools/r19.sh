#!/bin/bash
OUT=gpurun_out/r19; mkdir -p $OUT
for m in m3 m1qam16; do python tools/ofdm_quick_bench.py 4096 $m > $OUT/quick_$m.log 2>&1; tail -1 $OUT/quick_$m.log; done
( time python -m pytest tests -m gpu -x -q -k "ofdm" ) > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
