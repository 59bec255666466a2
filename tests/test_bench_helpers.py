"""CPU checks of the measurement plumbing: the roofline inputs bench.py reports are READ from the ncu summaries committed
under profiles/ (no constants in bench.py), the reference arm's config equals the GPU arm's, the golden proof of the headline
config is well-formed, and the worker-pool replay of the reference prover equals its serial replay."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import ref_wasm  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def test_roofline_inputs_come_from_the_committed_ncu_summaries():
    acc, src = bench.read_ncu_summary("k_msm_accumulate<G1Curve")
    assert src and src.startswith("profiles/") and os.path.exists(os.path.join(ROOT, src))
    traffic = acc["dram__bytes_read.sum"] + acc["dram__bytes_write.sum"]
    assert 1e8 < traffic < 1e10                                   # bytes per launch, units parsed (Mbyte / Gbyte)
    assert 0 < acc["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"] <= 100
    assert 0.1 < acc["gpu__time_duration.sum"] < 100              # ms
    ntt, src2 = bench.read_ncu_summary("k_ntt_pass")
    assert src2 and ntt["launch__registers_per_thread"] in (64, 80, 128)
    assert bench.read_ncu_summary("no_such_kernel") == ({}, None)
    src_text = open(os.path.join(ROOT, "bench.py")).read()
    assert "NCU_TRAFFIC_BYTES" not in src_text and "81.5" not in src_text      # round 1's hard-coded evidence is gone


def test_both_arms_name_the_same_metric_and_workload():
    assert bench.metric_name() == "Groth16 proofs/sec (BN128, 2^20 constraints)"
    cfg = bench.workload_config()
    assert cfg["log2_constraints"] == 20 and cfg["n_public"] == 58 and "2^20" in cfg["workload"]


def test_golden_proof_of_the_headline_config_is_committed():
    g = json.load(open(os.path.join(GOLD, "synth_2p20.golden.json")))
    assert g["log2_constraints"] == 20 and g["seed"] == 0xB204 and g["n_public"] == 58
    assert g["r"] + g["s"] == bench.RS.hex()                      # the (r, s) every timed proof uses
    assert len(g["pkey_sha256"]) == 64 and len(g["witness_sha256"]) == 64 and g["pkey_bytes"] == 501090524
    p = g["proof"]
    assert len(p["pi_a"]) == 3 and p["pi_a"][2] == "1" and len(p["pi_b"]) == 3 and p["pi_b"][2] == ["1", "0"] and p["pi_c"][2] == "1"


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not built")
def test_reference_worker_pool_replay_equals_the_serial_replay():
    """RefBn128Pool (CALC_H one task, every multiexp split over W host threads, H multiexp queued after CALC_H) computes the
    proof RefBn128 computes serially -- and the one stored for the small synthetic key."""
    pk = open(os.path.join(GOLD, "synth_small.pkey.bin"), "rb").read()
    wit = open(os.path.join(GOLD, "synth_small.witness.bin"), "rb").read()
    g = json.load(open(os.path.join(GOLD, "synth_small.golden.json")))
    r, s = bytes.fromhex(g["r"]), bytes.fromhex(g["s"])
    for threads in (1, 3):
        pool = ref_wasm.RefBn128Pool(concurrency=threads)
        try:
            assert pool.groth16GenProof(wit, pk, r=r, s=s) == g["proof_rs"]
        finally:
            pool.close()
    serial = ref_wasm.RefBn128(concurrency=3)
    try:
        assert serial.groth16GenProof(wit, pk, r=r, s=s) == g["proof_rs"]
    finally:
        serial.close()


def test_integer_roofline_reads_the_committed_census():
    """bench.py's `integer_roofline` object: the products of one 2^20-constraint proof as COUNTED by the kernel emulator
    (profiles/r2_op_census_2p20.json), turned into multiplier-pipe time; absent census -> None, never an error."""
    r = bench.integer_roofline(20, 17.41, 1965.0)
    census = json.load(open(os.path.join(ROOT, "profiles", "r2_op_census_2p20.json")))
    assert census["equals_reference_golden_proof"] is True and census["log2_constraints"] == 20
    assert r["products_per_proof"] == census["total"]["mul8"] and r["source"] == "profiles/r2_op_census_2p20.json"
    assert 7e8 < r["products_per_proof"] < 9e8 and 10.0 < r["pipe_ms_per_proof"] < 13.0 and 0.5 < r["frac"] < 0.8
    assert abs(bench.integer_roofline(20, 17.41, 982.5)["pipe_ms_per_proof"] - 2 * r["pipe_ms_per_proof"]) < 1e-9     # scales with the sampled clock
    assert bench.integer_roofline(7, 1.0) is None and bench.integer_roofline(20, 0.0)["frac"] is None
    # per-kernel: the accumulate kernels carry most of the products, one mixed addition = 8 products + 2 squarings on G1
    acc = census["kernels"]["k_msm_accumulate<G1Curve>"]
    assert acc["launches"] == 4 and acc["sqr8"] * 5 == acc["mul8"]
