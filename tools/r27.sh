#!/bin/bash
OUT=gpurun_out/r27; mkdir -p $OUT
( time timeout 1200 python -m pytest tests/test_chirp_sync_gpu.py -x -q ) > $OUT/pytest_chirp.log 2>&1; tail -6 $OUT/pytest_chirp.log | cut -c1-300
timeout 600 python tools/chirp_quick_bench.py 2048 > $OUT/chirp.log 2>&1; cat $OUT/chirp.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/bench.json 2> $OUT/bench.err; python -c "
import json
j=json.loads([l for l in open('gpurun_out/r27/bench.json') if l.startswith('{')][-1]); print(j['value'], j['e2e']['value'])"; tail -2 $OUT/bench.err
