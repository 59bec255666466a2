#!/bin/bash
# GPU session 2 (round 2, 2 GPUs): multi-GPU handle test, torchrun bench at N = 2, ws_multi timing.
set -x
O=gpurun_out/s2
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_prover.py -m gpu -x -q -k "multi or split or thread" > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log
tail -5 $O/pytest_multi.log
timeout 600 python tools/run_multi_once.py 20 1 10 > $O/multi_1.log 2>&1; tail -2 $O/multi_1.log
timeout 600 python tools/run_multi_once.py 20 2 10 > $O/multi_2.log 2>&1; tail -2 $O/multi_2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
tail -c 2000 $O/bench_n2.err
head -c 1500 $O/bench_n2.json
for np in 1 3 4; do WS_BENCH_PROOFS_INFLIGHT=$np timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --only-prove > $O/bench_n2_np$np.json 2> $O/bench_n2_np$np.err; done
ls -la $O
