#!/bin/bash
# GPU session 4 (round 2, 8 GPUs): the driver's scaling commands at N = 8 / 4 / 2, lane-group A/B, ws_multi on 8 GPUs.
set -x
O=gpurun_out/s4
mkdir -p $O
export WS_BENCH_KEY_CACHE=/tmp/wskeys
nvidia-smi -L > $O/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "n8 rc=$?"
tail -c 600 $O/bench_n8.err; head -c 600 $O/bench_n8.json
WS_BENCH_LANE_GROUPS=0 timeout 600 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --only-prove > $O/bench_n8_nogroups.json 2> $O/bench_n8_nogroups.err; echo "rc=$?"
WS_BENCH_PROOFS_INFLIGHT=6 timeout 600 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --steps 24 --warmup 6 --only-prove > $O/bench_n8_np6.json 2> $O/bench_n8_np6.err; echo "rc=$?"
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err; echo "n4 rc=$?" ) &
( CUDA_VISIBLE_DEVICES=4,5 timeout 900 $TR --nproc-per-node 2 --master-port 29525 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?" ) &
( CUDA_VISIBLE_DEVICES=6,7 timeout 600 python tools/run_multi_once.py 20 2 10 > $O/multi_2.log 2>&1 ) &
wait
timeout 600 python tools/run_multi_once.py 20 8 10 > $O/multi_8.log 2>&1; tail -1 $O/multi_8.log
timeout 300 python -m pytest tests/test_gpu_prover.py -m gpu -x -q -k "multi" > $O/pytest_multi8.log 2>&1; tail -2 $O/pytest_multi8.log
ls -la $O
