#!/bin/bash
OUT=gpurun_out/r25; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
tail -c 700 $OUT/bench_n$N.json | head -c 700; echo; tail -3 $OUT/bench_n$N.err | cut -c1-200
done
