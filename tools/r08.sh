#!/bin/bash
OUT=gpurun_out/r08; mkdir -p $OUT
( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
python tools/ofdm_quick_bench.py 4096 m1 > $OUT/quick_m1.log 2>&1; tail -1 $OUT/quick_m1.log
python tools/ofdm_quick_bench.py 4096 m3 > $OUT/quick_m3.log 2>&1; tail -1 $OUT/quick_m3.log
python tools/ofdm_quick_bench.py 4096 m1qam16 > $OUT/quick_m1qam16.log 2>&1; tail -1 $OUT/quick_m1qam16.log
