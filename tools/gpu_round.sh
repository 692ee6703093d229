#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full capture of the two hot kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
( time python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -3 $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref.json
if [ "$SKIP_NCU" != "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ofdm_ -s 3 -c 1 -f -o $OUT/prof_ofdm \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_ofdm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ldpc_flood -s 3 -c 1 -f -o $OUT/prof_ldpc \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/ncu_ldpc.log 2>&1
fi
ls -la $OUT
