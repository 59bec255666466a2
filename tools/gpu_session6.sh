#!/bin/bash
# GPU session 6 (round 2, 8 GPUs): balanced vs uniform split at N = 8 / 4 / 2, ws_multi on 8 GPUs.
set -x
O=gpurun_out/s6
mkdir -p $O
export WS_BENCH_KEY_CACHE=/tmp/wskeys
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "n8 rc=$?"
tail -c 400 $O/bench_n8.err; head -c 400 $O/bench_n8.json
WS_BENCH_SPLIT=uniform timeout 600 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 5 --only-prove > $O/bench_n8_uniform.json 2> $O/bench_n8_uniform.err; echo "rc=$?"
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 $TR --nproc-per-node 4 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err; echo "n4 rc=$?" ) &
( CUDA_VISIBLE_DEVICES=4,5 timeout 900 $TR --nproc-per-node 2 --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?" ) &
( CUDA_VISIBLE_DEVICES=6,7 WS_BENCH_SPLIT=balanced timeout 900 $TR --nproc-per-node 2 --master-port 29536 bench.py --gpus 2 --steps 20 --warmup 5 --only-prove > $O/bench_n2_balanced.json 2> $O/bench_n2_balanced.err; echo "n2b rc=$?" ) &
wait
( CUDA_VISIBLE_DEVICES=0,1,2,3 WS_BENCH_SPLIT=uniform timeout 600 $TR --nproc-per-node 4 --master-port 29537 bench.py --gpus 4 --steps 20 --warmup 5 --only-prove > $O/bench_n4_uniform.json 2> $O/bench_n4_uniform.err ) &
( CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 600 python tools/run_multi_once.py 20 4 10 > $O/multi_4.log 2>&1 ) &
wait
timeout 600 python tools/run_multi_once.py 20 8 10 > $O/multi_8.log 2>&1; tail -2 $O/multi_8.log
timeout 600 python tools/run_multi_once.py 20 8 10 0 > $O/multi_8_uniform.log 2>&1; tail -2 $O/multi_8_uniform.log
timeout 300 python -m pytest tests/test_gpu_prover.py -m gpu -x -q -k "multi" > $O/pytest_multi8.log 2>&1; tail -2 $O/pytest_multi8.log
ls -la $O
