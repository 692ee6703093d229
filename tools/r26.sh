#!/bin/bash
OUT=gpurun_out/r26; mkdir -p $OUT
( time timeout 1200 python -m pytest tests/test_chirp_sync_gpu.py -x -q ) > $OUT/pytest_chirp.log 2>&1; tail -30 $OUT/pytest_chirp.log | cut -c1-300
