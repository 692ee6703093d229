#!/bin/bash
OUT=gpurun_out/r40; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_linksim_gpu.py -x -q -m gpu ) > $OUT/pytest.log 2>&1; tail -15 $OUT/pytest.log | cut -c1-400
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; python -c "
import json; d=json.load(open('$OUT/bench.json')); print('windowed', d['value'], d['e2e'])"
PU_E2E_WHOLE_FRAMES=1 timeout 600 python bench.py > $OUT/bench_whole.json 2> $OUT/bench_whole.err; python -c "
import json; d=json.load(open('$OUT/bench_whole.json')); print('whole', d['value'], d['e2e'])"
