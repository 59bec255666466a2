#!/bin/bash
# GPU session 8 (round 2, 2 GPUs): the driver's scaling command at N = 2 with the final defaults (balanced split).
set -x
O=gpurun_out/s8
mkdir -p $O
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err ) 2> $O/bench_n2.time; echo "n2 rc=$?"; cat $O/bench_n2.time
tail -c 600 $O/bench_n2.err; head -c 600 $O/bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref rc=$?"; head -c 300 $O/bench_ref_n2.json
timeout 300 python -m pytest tests/test_gpu_prover.py -m gpu -x -q -k "multi" > $O/pytest_multi2.log 2>&1; tail -2 $O/pytest_multi2.log
timeout 300 python tools/run_multi_once.py 20 2 10 > $O/multi_2.log 2>&1; tail -2 $O/multi_2.log
