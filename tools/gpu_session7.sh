#!/bin/bash
# GPU session 7 (round 2, 1 GPU): final verification -- the driver's commands -- and the captures for profiles/.
set -x
O=gpurun_out/s7
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err ) 2> $O/bench_n1.time; echo "bench rc=$?"; cat $O/bench_n1.time
tail -c 400 $O/bench_n1.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/bench_ref.time; cat $O/bench_ref.time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_prove_2p20.csv python tools/run_prove_once.py 20 2 1 0 0 0 > $O/ncu_prove.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass16 -c 3 -o $O/ntt_pass16 python tools/run_ntt_once.py 22 1 0 > $O/ncu_ntt.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --only-prove > $O/ncu_bench.log 2>&1
ls -la $O
