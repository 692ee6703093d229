#!/bin/bash
OUT=gpurun_out/r17; mkdir -p $OUT
timeout 600 python tools/acquire_quick_bench.py 8192 > $OUT/acquire_quick.log 2>&1; cat $OUT/acquire_quick.log
ncu --set full --clock-control none --import-source on -k regex:ofdm_acquire -s 1 -c 1 -f -o $OUT/prof_acq python tools/acquire_quick_bench.py 2048 > $OUT/ncu_acq.log 2>&1
