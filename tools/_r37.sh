#!/bin/bash
OUT=gpurun_out/r37; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_psk_gpu.py tests/test_chirp_sync_gpu.py -x -q ) > $OUT/pytest.log 2>&1; tail -25 $OUT/pytest.log | cut -c1-300
