#!/bin/bash
# Sweep of the packed 512 kernel's variant switches (one process per variant: switches are read once).
OUT=${1:-gpurun_out/variants.log}
: > $OUT
PU_OFDM_NO_PACKED512=1 python tools/ofdm_quick_bench.py >> $OUT 2>&1
for half in 1 0; do for st in 2 3; do for w in 12 10 8 6 4; do
  PU_P512_THALF=$half PU_P512_STAGES=$st PU_P512_WARPS=$w python tools/ofdm_quick_bench.py >> $OUT 2>&1
done; done; done
cat $OUT
