#!/bin/bash
OUT=gpurun_out/r30; mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -x -q -k "chirp or dropin or ofdm_gpu" ) > $OUT/pytest.log 2>&1; tail -6 $OUT/pytest.log | cut -c1-300
timeout 600 python tools/chirp_quick_bench.py 2048 > $OUT/chirp.log 2>&1; cat $OUT/chirp.log
