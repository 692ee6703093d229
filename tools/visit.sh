#!/bin/bash
# Development GPU visit (round 2).  Usage: bash tools/visit.sh <tag> [what...]   what: tests fasttests quick bench workloads ncu
TAG=${1:-v01}; shift
WHAT=${@:-fasttests quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
for w in $WHAT; do case $w in
tests) ( time python -m pytest tests -m gpu -q -x ) > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log;;
fasttests) ( time python -m pytest tests/test_ofdm_fast_gpu.py -q -s ) > $OUT/pytest_fast.log 2>&1; grep -E "^mod|^fast|^against|^frames|passed|failed" $OUT/pytest_fast.log;;
quick) { python tools/ofdm_quick_bench.py 4096 m1; QB_PRECISION=fast python tools/ofdm_quick_bench.py 4096 m1; } > $OUT/quick.txt 2>&1; cat $OUT/quick.txt;;
smoke) python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log;;
bench) python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 6000 $OUT/bench.json; tail -5 $OUT/bench.err;;
benchref) python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref.json;;
workloads) for wl in ldpc m3 dpsk; do python bench.py --workload $wl --steps 5 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; tail -c 3000 $OUT/bench_$wl.json; tail -3 $OUT/bench_$wl.err; done;;
ncu)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --sustain-seconds 0 --sweep-batches 2 > $OUT/ncu_launch_bench.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:ofdm_fast -s 3 -c 1 -f -o $OUT/prof_ofdm \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --sustain-seconds 0 --sweep-batches 0 > $OUT/ncu_ofdm.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:ldpc_flood -s 3 -c 1 -f -o $OUT/prof_ldpc \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --sustain-seconds 0 --sweep-batches 0 > $OUT/ncu_ldpc.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:awgn_kernel -c 1 -f -o $OUT/prof_awgn \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --sustain-seconds 0 --sweep-batches 0 > $OUT/ncu_awgn.log 2>&1
  ;;
sweeptests) ( time python -m pytest tests/test_sweep_gpu.py -q -x ) > $OUT/pytest_sweep.log 2>&1; tail -15 $OUT/pytest_sweep.log;;
sweepsmoke) ( time projectultra_b200/pu_sweep --table smoke --trials 8192 --block 2048 --out $OUT/sweep_smoke.jsonl ) > $OUT/sweep_smoke.log 2>&1; tail -3 $OUT/sweep_smoke.log; tail -1 $OUT/sweep_smoke.jsonl;;
sweep2) for r in 0 1; do RANK=$r WORLD_SIZE=2 LOCAL_RANK=$r MASTER_PORT=29700 projectultra_b200/pu_sweep --table smoke --trials 8192 --block 1024 --rendezvous /tmp --out $OUT/sweep2.jsonl > $OUT/sweep2_r$r.log 2>&1 & done; wait; tail -2 $OUT/sweep2_r0.log $OUT/sweep2_r1.log; tail -1 $OUT/sweep2.jsonl;;
bench2) python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 > $OUT/bench2.json 2> $OUT/bench2.err; tail -c 2500 $OUT/bench2.json; tail -3 $OUT/bench2.err;;
chantests) ( time python -m pytest tests/test_channel.py tests/test_linksim_gpu.py tests/test_sweep_gpu.py -q -x ) > $OUT/pytest_chan.log 2>&1; tail -6 $OUT/pytest_chan.log;;
dropin) ( time python -m pytest tests/test_dropin_cpp_gpu.py tests/test_ofdm_gpu.py -q -x ) > $OUT/pytest_dropin.log 2>&1; tail -30 $OUT/pytest_dropin.log; oracle/_ref/dropin_driver > $OUT/dropin_driver.log 2>&1; tail -30 $OUT/dropin_driver.log; oracle/_ref/test_multiblock_ldpc_pu > $OUT/multiblock.log 2>&1; tail -12 $OUT/multiblock.log;;
ncu3)
  for m in m3 m1qam16; do ncu --set full --clock-control none --import-source on -k regex:ofdm_presynced -s 3 -c 1 -f -o $OUT/prof_$m python tools/ofdm_quick_bench.py 4096 $m > $OUT/ncu_$m.log 2>&1; done
  QB_CHANNEL=good ncu --set full --clock-control none --import-source on -k regex:channel_kernel -c 1 -f -o $OUT/prof_channel python tools/ofdm_quick_bench.py 4096 m1 > $OUT/ncu_channel.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:awgn_kernel -c 1 -f -o $OUT/prof_awgn python tools/ofdm_quick_bench.py 4096 m1 > $OUT/ncu_awgn.log 2>&1
  for m in m3 m1qam16; do python tools/ofdm_quick_bench.py 4096 $m; done > $OUT/quick3.txt 2>&1; cat $OUT/quick3.txt;;
sanitize)
  # compute-sanitizer over a reduced subset that launches every kernel family; PU_SANITIZE=1 shrinks the batches
  export PU_SANITIZE=1
  for tool in memcheck racecheck; do
    ( time compute-sanitizer --tool $tool --error-exitcode 9 --log-file $OUT/sanitizer_$tool.log python -m pytest tests/test_sanitize_gpu.py -q -x ) > $OUT/sanitize_$tool.out 2>&1
    echo "$tool rc=$?"; tail -3 $OUT/sanitize_$tool.out; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/sanitizer_$tool.log | tail -2
  done;;
psktx) ( time python -m pytest tests/test_psk_tx_gpu.py tests/test_ldpc_gpu.py -q -x ) > $OUT/pytest_psktx.log 2>&1; tail -25 $OUT/pytest_psktx.log;;
fer) ( time python -m pytest tests/test_fer_curves_gpu.py -q -x -s ) > $OUT/pytest_fer.log 2>&1; grep -E "dB|passed|failed|Error" $OUT/pytest_fer.log | tail -30;;
sweepred) ( time projectultra_b200/pu_sweep --table reduced --trials ${TRIALS:-4096} --block 1024 --out $OUT/sweep_reduced.jsonl ) > $OUT/sweep_reduced.log 2>&1; tail -4 $OUT/sweep_reduced.log; tail -1 $OUT/sweep_reduced.jsonl;;
sweepN) N=${NGPU:-8}; for r in $(seq 0 $((N-1))); do RANK=$r WORLD_SIZE=$N LOCAL_RANK=$r MASTER_PORT=29700 projectultra_b200/pu_sweep --table reduced --trials ${TRIALS:-16384} --block 1024 --rendezvous /tmp --manifest $OUT/manifest_N$N --out $OUT/sweep_N$N.jsonl > $OUT/sweepN_r$r.log 2>&1 & done; wait; tail -n 2 $OUT/sweepN_r0.log; tail -1 $OUT/sweep_N$N.jsonl;;
h2dN) N=${NGPU:-8}; for n in 1 2 4 8; do [ $n -le $N ] && python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2972$n tools/h2d_bench.py 2>/dev/null | grep '^{' | tee -a $OUT/h2d.jsonl | cut -c1-400; done;;
benchN) N=${NGPU:-8}; for n in ${BENCH_NS:-4 8}; do [ $n -le $N ] && python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2973$n bench.py --gpus $n 2>$OUT/bench_N$n.err | grep '^{' > $OUT/bench_N$n.json; python -c "import json,sys; j=json.load(open('$OUT/bench_N$n.json')); print($n, 'value', j['value'], 'e2e', j['e2e']['value'], j['e2e'].get('h2d_gbs_per_gpu'), 'sweep', {k:v['value'] for k,v in j.get('sweep',{}).items()})"; done;;
frames) ( time python -m pytest tests/test_frame_gpu.py -q -x ) > $OUT/pytest_frames.log 2>&1; tail -25 $OUT/pytest_frames.log;;
fastgen) ( time python -m pytest tests/test_ofdm_fast_gpu.py -q -s -k "general or falls" ) > $OUT/pytest_fastgen.log 2>&1; grep -E "^case|passed|failed|Error|assert" $OUT/pytest_fastgen.log | cut -c1-400 | tail -30
  { for m in m3 m1qam16; do python tools/ofdm_quick_bench.py 4096 $m; QB_PRECISION=fast python tools/ofdm_quick_bench.py 4096 $m; done; QB_CHANNEL=good QB_PRECISION=fast python tools/ofdm_quick_bench.py 4096 m3; } > $OUT/quick_fastgen.txt 2>&1; cat $OUT/quick_fastgen.txt;;
chirp) ( time python -m pytest tests/test_chirp_sync_gpu.py tests/test_psk_gpu.py tests/test_dropin_cpp_gpu.py -q -s -x ) > $OUT/pytest_chirp.log 2>&1; grep -E "two-tier|passed|failed|Error|assert" $OUT/pytest_chirp.log | cut -c1-300 | tail -12
  { python tools/chirp_quick_bench.py 2048; python tools/chirp_quick_bench.py 2048 mcdpsk; PU_CHIRP_SEARCH=exact python tools/chirp_quick_bench.py 1024; } > $OUT/quick_chirp.txt 2>&1; grep -v "^ *$" $OUT/quick_chirp.txt | cut -c1-220;;
ncu5)
  ncu --set full --clock-control none --import-source on -k regex:chirp_detect -c 1 -f -o $OUT/prof_chirp2 python tools/chirp_quick_bench.py 1024 > $OUT/ncu_chirp2.log 2>&1;;
ncu6)
  ncu --set full --clock-control none --import-source on -k regex:dpsk_find_preamble -c 1 -f -o $OUT/prof_bk python tools/dpsk_acquire_quick_bench.py 1024 > $OUT/ncu_bk.log 2>&1; tail -5 $OUT/ncu_bk.log
  ncu --set full --clock-control none --import-source on -k regex:ofdm_acquire_kernel -c 1 -f -o $OUT/prof_acq python tools/acquire_quick_bench.py 1024 > $OUT/ncu_acq.log 2>&1; tail -5 $OUT/ncu_acq.log;;
sweep5) ( time projectultra_b200/pu_sweep --table config5 --trials ${TRIALS:-1024} --block 1024 --out $OUT/sweep_config5.jsonl ) > $OUT/sweep_config5.log 2>&1; tail -6 $OUT/sweep_config5.log; tail -1 $OUT/sweep_config5.jsonl | cut -c1-600;;
sweep5N) N=${NGPU:-8}; for r in $(seq 0 $((N-1))); do RANK=$r WORLD_SIZE=$N LOCAL_RANK=$r MASTER_PORT=29700 projectultra_b200/pu_sweep --table config5 --trials ${TRIALS:-8192} --block 1024 --rendezvous /tmp --manifest $OUT/manifest5_N$N --out $OUT/sweep_config5_N$N.jsonl > $OUT/sweep5N_r$r.log 2>&1 & done; wait; tail -n 3 $OUT/sweep5N_r0.log; tail -1 $OUT/sweep_config5_N$N.jsonl | cut -c1-800;;
ncu4)
  for m in m3 m1qam16; do QB_PRECISION=fast ncu --set full --clock-control none --import-source on -k regex:ofdm_presynced -s 3 -c 1 -f -o $OUT/prof_fast_$m python tools/ofdm_quick_bench.py 4096 $m > $OUT/ncu_fast_$m.log 2>&1; done;;
*) echo "unknown: $w";;
esac; done
ls -la $OUT
