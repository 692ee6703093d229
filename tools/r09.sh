#!/bin/bash
OUT=gpurun_out/r09; mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:ofdm_presynced -s 2 -c 1 -f -o $OUT/prof_m3 python tools/ofdm_quick_bench.py 4096 m3 > $OUT/ncu_m3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ofdm_diff512 -s 2 -c 1 -f -o $OUT/prof_ofdm python tools/ofdm_quick_bench.py 4096 m1 > $OUT/ncu_ofdm.log 2>&1
V=$OUT/variants.log; : > $V
for w in 16 14 12; do PU_P512_INPLACE=1 PU_P512_WARPS=$w python tools/ofdm_quick_bench.py >> $V 2>&1; done
PU_P512_STAGES=3 python tools/ofdm_quick_bench.py >> $V 2>&1
python tools/ofdm_quick_bench.py >> $V 2>&1
cat $V
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/f32x2_bench tools/ubench/f32x2_bench.cu && /tmp/f32x2_bench > $OUT/f32x2.log 2>&1; cat $OUT/f32x2.log
