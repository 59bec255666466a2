#!/bin/bash
# GPU session 3 (round 2, 1 GPU): tests with the cooperative reduction levels + warp-local NTT exchange, timings, ncu.
set -x
O=gpurun_out/s3
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for v in 0 3 4; do timeout 300 python tools/run_ntt_once.py 22 4 $v; timeout 300 python tools/run_ntt_once.py 20 4 $v; done > $O/ntt_variants.log 2>&1
cat $O/ntt_variants.log
timeout 300 python tools/run_msm_once.py 20 4 > $O/msm_once.log 2>&1; tail -1 $O/msm_once.log
timeout 300 python tools/run_prove_once.py 20 3 1 0 0 10 > $O/prove_once.log 2>&1; tail -3 $O/prove_once.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
tail -c 800 $O/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_prove_2p20.csv python tools/run_prove_once.py 20 2 1 0 0 0 > $O/ncu_prove.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate.*G1Curve -c 2 -o $O/acc_g1_table python tools/run_prove_once.py 20 1 1 0 0 0 > $O/ncu_acc.log 2>&1
for v in 3; do timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass64 -c 2 -o $O/ntt_pass64w python tools/run_ntt_once.py 22 1 $v > $O/ncu_ntt.log 2>&1; done
ls -la $O
