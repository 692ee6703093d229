#!/bin/bash
OUT=gpurun_out/r05; mkdir -p $OUT
( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
V=$OUT/variants.log; : > $V
PU_P512_INPLACE=0 python tools/ofdm_quick_bench.py >> $V 2>&1
for w in 16 15 14 13 12; do PU_P512_INPLACE=1 PU_P512_WARPS=$w python tools/ofdm_quick_bench.py >> $V 2>&1; done
cat $V
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $OUT/bench_a.json 2> $OUT/bench_a.err
PU_LDPC_MINB4=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $OUT/bench_minb4.json 2> $OUT/bench_minb4.err
python - <<'PY'
import json
for f in ("bench_a","bench_minb4"):
    try:
        j=json.loads([l for l in open("gpurun_out/r05/%s.json"%f) if l.startswith("{")][-1])
        print(f, j["value"], j["stages_ms"], j["roofline"]["frac"], j["e2e"]["value"])
    except Exception as e: print(f, "failed", e)
PY
