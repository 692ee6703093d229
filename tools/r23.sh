#!/bin/bash
OUT=gpurun_out/r23; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_dpsk_acquire_gpu.py tests/test_psk_gpu.py -x -q ) > $OUT/pytest_dpsk.log 2>&1; tail -30 $OUT/pytest_dpsk.log | cut -c1-300
