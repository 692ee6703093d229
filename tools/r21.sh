#!/bin/bash
OUT=gpurun_out/r21; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; tail -25 $OUT/pytest_gpu.log | cut -c1-300
