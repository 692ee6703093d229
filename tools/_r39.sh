#!/bin/bash
OUT=gpurun_out/r39; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_frame_gpu.py -x -q -m gpu ) > $OUT/pytest.log 2>&1; tail -30 $OUT/pytest.log | cut -c1-500
