#!/bin/bash
OUT=gpurun_out/r20; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_ofdm_tx_gpu.py -x -q ) > $OUT/pytest_tx.log 2>&1; tail -25 $OUT/pytest_tx.log | cut -c1-300
