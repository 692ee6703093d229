#!/bin/bash
OUT=gpurun_out/r06; mkdir -p $OUT
( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $OUT/bench_a.json 2> $OUT/bench_a.err
PU_P512_INPLACE=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $OUT/bench_b.json 2> $OUT/bench_b.err
python - <<'PY'
import json
for f in ("bench_a","bench_b"):
    try:
        j=json.loads([l for l in open("gpurun_out/r06/%s.json"%f) if l.startswith("{")][-1])
        print(f, j["value"], j["stages_ms"], j["roofline"]["frac"], j["e2e"]["value"])
    except Exception as e: print(f, "failed", e)
PY
python tools/ldpc_quick_bench.py 262144 > $OUT/ldpc_quick.log 2>&1; cat $OUT/ldpc_quick.log
