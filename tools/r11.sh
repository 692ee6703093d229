#!/bin/bash
OUT=gpurun_out/r11; mkdir -p $OUT
V=$OUT/variants.log; : > $V
python tools/ofdm_quick_bench.py >> $V 2>&1
PU_P512_FASTSTORE=0 python tools/ofdm_quick_bench.py >> $V 2>&1
python tools/ofdm_quick_bench.py >> $V 2>&1
PU_P512_FASTSTORE=0 python tools/ofdm_quick_bench.py >> $V 2>&1
cat $V
python tools/ldpc_quick_bench.py 262144 > $OUT/ldpc_quick.log 2>&1; grep "rate=2" $OUT/ldpc_quick.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $OUT/bench_a.json 2> $OUT/bench_a.err
python -c "
import json
j=json.loads([l for l in open('gpurun_out/r11/bench_a.json') if l.startswith('{')][-1])
print(j['value'], j['stages_ms'], j['roofline']['frac'], j['e2e']['value'])"
ncu --set full --clock-control none --import-source on -k regex:ofdm_presynced -s 2 -c 1 -f -o $OUT/prof_m3w python tools/ofdm_quick_bench.py 4096 m3 > $OUT/ncu_m3.log 2>&1
( time python -m pytest tests -m gpu -x -q -k "ldpc or linksim" ) > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
