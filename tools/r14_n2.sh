#!/bin/bash
OUT=gpurun_out/r14; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
tail -c 1800 $OUT/bench_n2.json; tail -5 $OUT/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/bench_ref_n2.json 2> $OUT/bench_ref_n2.err
tail -c 400 $OUT/bench_ref_n2.json
