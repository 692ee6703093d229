#!/bin/bash
OUT=gpurun_out/r18; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q -k "acquire or linksim or dropin" ) > $OUT/pytest_gpu.log 2>&1; tail -12 $OUT/pytest_gpu.log | cut -c1-300
timeout 600 python tools/acquire_quick_bench.py 8192 > $OUT/acquire_quick.log 2>&1; cat $OUT/acquire_quick.log
