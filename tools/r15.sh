#!/bin/bash
OUT=gpurun_out/r15; mkdir -p $OUT
( time timeout 600 python -m pytest tests/test_ofdm_acquire_gpu.py -x -q ) > $OUT/pytest_acq.log 2>&1; tail -30 $OUT/pytest_acq.log | cut -c1-300
