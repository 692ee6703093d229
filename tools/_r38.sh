#!/bin/bash
OUT=gpurun_out/r38; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_dropin_cpp_gpu.py tests/test_psk_gpu.py -x -q -m gpu ) > $OUT/pytest.log 2>&1; tail -25 $OUT/pytest.log | cut -c1-400
timeout 600 python tools/chirp_quick_bench.py 2048 mcdpsk > $OUT/chirp_mcdpsk.log 2>&1; cat $OUT/chirp_mcdpsk.log | tail
