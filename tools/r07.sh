#!/bin/bash
bash tools/gpu_round.sh r07 > gpurun_out/r07_round.log 2>&1
OUT=gpurun_out/r07
python tools/ofdm_quick_bench.py 4096 m3 > $OUT/quick_m3.log 2>&1; cat $OUT/quick_m3.log | tail -2
python tools/ofdm_quick_bench.py 4096 m1qam16 > $OUT/quick_m1qam16.log 2>&1; cat $OUT/quick_m1qam16.log | tail -2
QB_CHANNEL=good python tools/ofdm_quick_bench.py 4096 m3 > $OUT/quick_m3_good.log 2>&1; cat $OUT/quick_m3_good.log | tail -2
ncu --set full --clock-control none --import-source on -k regex:ofdm_presynced -s 2 -c 1 -f -o $OUT/prof_m3 python tools/ofdm_quick_bench.py 4096 m3 > $OUT/ncu_m3.log 2>&1
tail -3 $OUT/pytest_gpu.log; tail -c 1500 $OUT/bench.json
