#!/bin/bash
# GPU session 10 (round 2, 1 GPU): final state -- tests, smoke, the bench command under ncu (launch list of every launch).
set -x
O=gpurun_out/s10
mkdir -p $O
export WS_BENCH_KEY_CACHE=/tmp/wskeys
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; head -c 300 $O/bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --only-prove > $O/ncu_bench.log 2>&1; tail -2 $O/ncu_bench.log; wc -l $O/launches_bench.csv
