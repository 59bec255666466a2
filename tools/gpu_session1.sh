#!/bin/bash
# GPU session 1 (round 2): tests, N=1 bench, NTT variants, launch lists.  Everything lands in gpurun_out/s1/.
set -x
O=gpurun_out/s1
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for v in 0 1 2; do timeout 300 python tools/run_ntt_once.py 22 4 $v; timeout 300 python tools/run_ntt_once.py 20 4 $v; done > $O/ntt_variants.log 2>&1
cat $O/ntt_variants.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
tail -c 1500 $O/bench_n1.err
head -c 3000 $O/bench_n1.json
WS_BENCH_PROOFS_INFLIGHT=3 timeout 600 python bench.py --steps 10 --warmup 3 --only-prove > $O/bench_n1_np3.json 2> $O/bench_n1_np3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_prove_2p20.csv python tools/run_prove_once.py 20 2 1 0 0 0 > $O/ncu_prove.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass64 -c 3 -o $O/ntt_pass64 python tools/run_ntt_once.py 22 1 0 > $O/ncu_ntt.log 2>&1
ls -la $O
