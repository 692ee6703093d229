#!/bin/bash
OUT=gpurun_out/r24; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q -k "linksim or dpsk or psk" ) > $OUT/pytest_gpu.log 2>&1; tail -12 $OUT/pytest_gpu.log | cut -c1-300
timeout 600 python tools/dpsk_acquire_quick_bench.py 2048 > $OUT/dpsk_acquire.log 2>&1; cat $OUT/dpsk_acquire.log
