#!/bin/bash
# GPU session 9 (round 2, 1 GPU): one sweep of the tuning hooks on the whole proof (proofs/s, --only-prove).
O=gpurun_out/s9
mkdir -p $O
export WS_BENCH_KEY_CACHE=/tmp/wskeys
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 12 --warmup 3 --only-prove > $O/$name.json 2> $O/$name.err; python - <<PY
import json
try:
    b=json.load(open("$O/$name.json")); print("$name", "value %.2f lat %.2f e2e %.2f" % (b["value"], b["latency"]["ms_per_proof_single_stream"], b["e2e"]["value"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
run base X=1
run seg4 WS_MSM_SEG_LEN=4
run seg16 WS_MSM_SEG_LEN=16
run g2v1 WS_MSM_ACC_VARIANT=16
run g2v2 WS_MSM_ACC_VARIANT=32
run g1v3 WS_MSM_ACC_VARIANT=3
run np4 WS_BENCH_PROOFS_INFLIGHT=4
run base2 X=1
