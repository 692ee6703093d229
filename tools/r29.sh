#!/bin/bash
OUT=gpurun_out/r29; mkdir -p $OUT
for m in m3 m1qam16; do
  python tools/ofdm_quick_bench.py 4096 $m > $OUT/quick_$m.log 2>&1; tail -1 $OUT/quick_$m.log
  PU_OFDM_WARPG_SYNC=1 python tools/ofdm_quick_bench.py 4096 $m > $OUT/quick_${m}_sync.log 2>&1; tail -1 $OUT/quick_${m}_sync.log
done
( time PU_OFDM_WARPG_SYNC=1 python -m pytest tests -m gpu -x -q -k "ofdm or acquire or chirp" ) > $OUT/pytest_sync.log 2>&1; tail -3 $OUT/pytest_sync.log
