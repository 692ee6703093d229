#!/bin/bash
OUT=gpurun_out/r12; mkdir -p $OUT
for m in m1 m3 m1qam16; do python tools/ofdm_quick_bench.py 4096 $m > $OUT/quick_$m.log 2>&1; tail -1 $OUT/quick_$m.log; done
( time python -m pytest tests -m gpu -x -q -k "ofdm" ) > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
ncu --set full --clock-control none --import-source on -k regex:ofdm_presynced -s 2 -c 1 -f -o $OUT/prof_m3w python tools/ofdm_quick_bench.py 4096 m3 > $OUT/ncu_m3.log 2>&1
