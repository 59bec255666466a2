#!/bin/bash
# GPU session 5 (round 2, 1 GPU): balanced-split tests, radix-16 NTT, table window size A/B, ncu of the accumulate kernel.
set -x
O=gpurun_out/s5
mkdir -p $O
export WS_BENCH_KEY_CACHE=/tmp/wskeys
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for v in 0 4 5 6; do timeout 300 python tools/run_ntt_once.py 22 4 $v; timeout 300 python tools/run_ntt_once.py 20 4 $v; done > $O/ntt_variants.log 2>&1
grep variant $O/ntt_variants.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
tail -c 500 $O/bench_n1.err
WS_TABLE_C=19 timeout 600 python bench.py --steps 10 --warmup 3 --only-prove > $O/bench_n1_c19.json 2> $O/bench_n1_c19.err
WS_TABLE_C=18 timeout 600 python bench.py --steps 10 --warmup 3 --only-prove > $O/bench_n1_c18.json 2> $O/bench_n1_c18.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -c 5 -o $O/acc_table python tools/run_prove_once.py 20 1 1 0 0 0 > $O/ncu_acc.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_prove_2p20.csv python tools/run_prove_once.py 20 2 1 0 0 0 > $O/ncu_prove.log 2>&1
ls -la $O
