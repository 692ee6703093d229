#!/bin/bash
OUT=gpurun_out/r35; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_linksim_gpu.py -x -q ) > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:ofdm_presynced -s 2 -c 1 -f -o $OUT/prof_m3s python tools/ofdm_quick_bench.py 4096 m3 > $OUT/ncu_m3.log 2>&1
