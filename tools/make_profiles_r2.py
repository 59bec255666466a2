"""Writes profiles/README.md from the round-2 artefacts under profiles/ (bench lines, launch lists, ncu summaries).
usage: python tools/make_profiles_r2.py"""
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    try:
        return json.load(open(os.path.join(P, name)))
    except Exception:
        return None


ONE_PROOF_SHARE = {}


def launch_table(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    L = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[start:] if len(r) > mv]
    idx = [i for i, (k, _) in enumerate(L) if "k_assemble_post" in k]
    if len(idx) >= 2:
        L = L[idx[-2] + 1:idx[-1] + 1]            # the launches of the last proof
    a = collections.OrderedDict()
    for k, v in L:
        k = re.sub(r"\(.*", "", k).replace("void ", "").replace("ws::", "")
        a.setdefault(k, [0, 0.0])
        a[k][0] += 1
        a[k][1] += v
    tot = sum(v for _, v in a.values())
    ONE_PROOF_SHARE["acc_g1"] = 100 * sum(v for k, (_, v) in a.items() if k.startswith("k_msm_accumulate<G1Curve")) / tot
    lines = ["| `%s` | %d | %.1f | %.1f %% |" % (k[:60], c, v / 1000.0, 100 * v / tot) for k, (c, v) in sorted(a.items(), key=lambda kv: -kv[1][1])]
    return lines, tot / 1e6


def main():
    out = ["# profiles -- round 2 (B200, sm_100a)", ""]
    b1 = load("r2_bench_n1.json")
    ref = load("r2_bench_n1_reference_arm.json")
    out += ["Files (round 1 files keep their `r1_` prefix and are described in the git history of this README):",
            "* `r2_bench_n1.json` -- `python bench.py --steps 10 --warmup 3` (N = 1); `r2_bench_n1_reference_arm.json` -- `--impl reference`.",
            "* `r2_bench_n{2,4,8}.json` -- the same command under torchrun (NCCL), one rank per GPU: the strong-scaling line.",
            "* `r2_bench_n8_only_prove_*.json`, `r2_bench_n2_only_prove_*.json` -- A/B runs (`--only-prove`): proofs in flight, one NCCL communicator vs one per lane, uniform vs balanced split.",
            "* `r2_launches_prove_2p20.csv` -- `ncu --metrics gpu__time_duration.sum --clock-control none` launch list of `tools/run_prove_once.py 20 2` (cold-cache, serialised: compare SHARES).",
            "* `r2_ncu_accumulate_*.txt`, `r2_ncu_ntt_*.txt` -- `ncu --set full --clock-control none` summaries (tools/ncu_summary.py) of the accumulate kernels of one proof and of the NTT passes.",
            "* `r2_ntt_variants*.txt` -- `tools/run_ntt_once.py` timings of the NTT pass variants (occupancy, warp-local exchange, radix 16).",
            "* `r2_op_census_2p20.json` / `.md` -- Montgomery products and squarings of one 2^20-constraint proof, counted per kernel inside the executed kernels (CPU kernel emulator, `tests/host/cuemu/census.py`), next to the measured kernel times.",
            "* `r2_device_sass.sha256` -- hash of the SASS of all 177 kernels of the build the last GPU session ran; `r2_device_sass_per_kernel.txt` -- the same per kernel for the CURRENT build, each kernel marked `gpu-verified` (byte-identical to that build: 171 kernels) or `changed-after-last-gpu-run` (6 kernels, see the note at the end); `tools/sass_hash.py` writes both, `tests/test_abi.py` checks the built library against them.",
            "* `r2_ptxas_resources.txt` -- registers / stack / spill bytes of every kernel (`-Xptxas -v`); `r2_tuning_hook_sweep.txt`, `r2_ws_multi_timings.txt` -- see the A/B table at the end.", ""]
    if b1:
        cb = b1.get("cpu_baseline") or {}
        out += ["## Headline: Groth16 proofs/s, BN128, 2^20 constraints (one B200; clocks %s)" % json.dumps(b1.get("clocks")), "",
                "| quantity | value |", "|---|---|",
                "| value (key + witness resident, %s) | %.1f proofs/s (%.2f ms per proof) |" % (b1["config"]["in_flight"], b1["value"], b1["ms_per_step"]),
                "| single-proof latency | %.2f ms |" % b1["latency"]["ms_per_proof_single_stream"],
                "| e2e: blocking host-buffer call, pinned witness | %.1f proofs/s; pageable witness %.1f |" % (b1["e2e"]["value"], b1["e2e"].get("pageable", {}).get("value", float("nan"))),
                "| the stateless drop-in call `ws_bn128_groth16_prove(witness, pkey)` | first call %.2f s, then %.1f ms per call |" % (b1["dropin_call"]["first_call_s"], b1["dropin_call"]["repeat_call_ms"]) if b1.get("dropin_call") else "| drop-in call | n/a |",
                "| bit-exact vs the reference's 2^20 proof (tests/golden) | %s |" % b1["config"].get("bit_exact_vs_reference_golden"),
                "| CPU baseline: the reference's groth16GenProof, %s threads | %.4f proofs/s (single thread %.5f) -- %s |" % (cb.get("cores"), cb.get("value") or 0, cb.get("single_thread_value") or 0, (cb.get("sample") or "")[:160]),
                ]
        rf = b1.get("roofline") or {}
        if rf:
            out += ["| dominant kernel %s | %.3f ms per launch; %d algorithmic B => %.1f GB/s = %.2f %% of the measured %.1f GB/s; ncu: fmaheavy pipe %s %%, DRAM traffic %s B per launch (%s) |"
                    % (rf["kernel"], rf["kernel_ms"], rf["algorithmic_bytes_per_launch"], rf["achieved"], 100 * rf["frac"], rf["peak"], (rf.get("ncu") or {}).get("fmaheavy_pipe_busy_pct"), rf.get("traffic"), rf.get("traffic_source"))]
        out += ["", "Stage stamps of one proof (CUDA events, ms): " + ", ".join("%s %.3f" % kv for kv in (b1.get("stages_ms") or {}).items()),
                "(" + (b1.get("stages_note") or "stamps on the H stream only: the witness MSMs are forked onto side streams before the H chain and fill the SMs first, so 'h:ntt' is the H chain's wall time while time-sliced with the witness MSMs (about 2.3 ms alone), not NTT kernel time") + ")", ""]
        m = b1.get("msm") or {}
        if m:
            out += ["## G1 MSM 2^20 (BASELINE configs[1])", "", "| quantity | value |", "|---|---|",
                    "| value (4 MSMs in flight) | %.1f M pairs/s (%.2f ms) |" % (m["value"] / 1e6, m["ms_per_step"]),
                    "| single-stream latency | %.2f ms |" % m["latency"]["ms_per_msm_single_stream"],
                    "| e2e (host buffers) | %.1f M pairs/s |" % (m["e2e"]["value"] / 1e6),
                    "| reference `g1m_multiexp2`, %s threads | %.0f pairs/s |" % ((m.get("cpu_baseline") or {}).get("cores"), (m.get("cpu_baseline") or {}).get("value") or 0),
                    "| stages (ms) | " + ", ".join("%s %.3f" % kv for kv in m["stages_ms"].items()) + " |", ""]
        o = ["## Other lines of the same run", "", "| item | value |", "|---|---|"]
        if b1.get("ntt"):
            o.append("| NTT 2^22 forward + inverse (configs[2]) | %.3f ms => %.0f GB/s = %.1f %% of HBM peak |" % (b1["ntt"]["ms_per_round_trip"], b1["ntt"]["roofline"]["achieved"], 100 * b1["ntt"]["roofline"]["frac"]))
        if b1.get("example_key"):
            e = b1["example_key"]
            o.append("| example/bn128 key (configs[3]) | %.2f ms per proof, equal to proof_good0.json: %s |" % (e["ms_per_proof_single_stream"], e["bit_exact_vs_proof_good0"]))
        if b1.get("mnt6753"):
            o.append("| MNT6753 G1 MSM 2^18 (configs[4]) | %.1f ms, closed form %s |" % (b1["mnt6753"]["ms_per_msm"], b1["mnt6753"]["matches_closed_form"]))
        if b1.get("bls12381"):
            o.append("| BLS12-381 G1 MSM 2^20 | %.1f ms, closed form %s |" % (b1["bls12381"]["ms_per_msm"], b1["bls12381"]["matches_closed_form"]))
        if b1.get("verifier"):
            v = b1["verifier"]
            o.append("| host verifier | %.1f ms (reference image %.1f ms) |" % (v["ms_per_verify"], v.get("reference_ms_per_verify") or float("nan")))
        out += o + [""]
    if ref:
        out += ["Reference arm (`--impl reference`): %.4f proofs/s -- %s" % (ref["value"], ref["cpu_baseline"]["sample"]), ""]
    rows = []
    base = b1["value"] if b1 else None
    for n in (1, 2, 4, 8):
        d = load("r2_bench_n%d.json" % n)
        if not d:
            continue
        m = d.get("msm") or {}
        mw = d.get("msm_weak") or {}
        rows.append("| %d | %.1f | %.2f | %.2f | %.1f | %s | %s | %s | %s |" % (
            n, d["value"], d["ms_per_step"], d["latency"]["ms_per_proof_single_stream"], d["e2e"]["value"],
            "%.2f" % (d["value"] / (n * base)) if base else "-",
            "%.0f" % (m["value"] / 1e6) if m else "-", "%.2f" % m["latency"]["ms_per_msm_single_stream"] if m else "-",
            "%.0f" % (mw["value"] / 1e6) if mw else "-"))
    if rows:
        out += ["## Strong scaling (torchrun, one rank per GPU, NCCL): one fixed 2^20-constraint proof / one fixed 2^20-pair MSM", "",
                "| GPUs | proofs/s | ms per proof | single-proof latency ms | e2e proofs/s | efficiency vs N = 1 | MSM 2^20 M pairs/s (strong) | MSM latency ms | MSM weak M pairs/s |", "|---|---|---|---|---|---|---|---|---|"] + rows + [""]
    lp = os.path.join(P, "r2_launches_prove_2p20.csv")
    if os.path.exists(lp):
        t, tot = launch_table(lp)
        out += ["## Launch list shares -- one Groth16 proof, 2^20 constraints, window tables (%.1f ms serialised under ncu)" % tot, "", "| kernel | launches | us | share |", "|---|---|---|---|"] + t + [""]
    lb = os.path.join(P, "r2_launches_bench_only_prove.csv")
    if os.path.exists(lb):
        rows = list(csv.reader(open(lb)))
        for i, r in enumerate(rows):
            if "Kernel Name" in r:
                hdr, start = r, i + 1
                break
        kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        a = collections.OrderedDict()
        for r in rows[start:]:
            if len(r) <= mv:
                continue
            k = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("ws::", "")
            a.setdefault(k, [0, 0.0])
            a[k][0] += 1
            a[k][1] += float(r[mv].replace(",", ""))
        tot = sum(v for _, v in a.values())
        # the proving kernels: key-load work (window / comb tables, twiddle table) happens once per key, outside every timed region
        lib = [(k, c, v) for k, (c, v) in a.items() if k.startswith("k_") and not k.startswith(("k_msm_build_table", "k_build_fixed_tables", "k_tw_"))]
        lib_tot = sum(v for _, _, v in lib)
        out += ["## Launch list of the bench command itself (`r2_launches_bench_only_prove.csv`)", "",
                "`ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 python bench.py --steps 2 --warmup 3 --only-prove` (every launch of the run: the example-key proofs, "
                "the 2^20 proofs of all legs, the table-MSM loops of the roofline leg; the library's proving kernels are %.1f %% of the captured time, the rest is key-load work and torch fills / copies of the harness):" % (100 * lib_tot / tot), "",
                "| kernel | launches | us | share of the proving kernels |", "|---|---|---|---|"]
        out += ["| `%s` | %d | %.1f | %.1f %% |" % (k[:60], c, v / 1000.0, 100 * v / lib_tot) for k, c, v in sorted(lib, key=lambda x: -x[2])[:14]] + [""]
        acc = sum(v for k, _, v in lib if k.startswith("k_msm_accumulate<G1Curve"))
        out += ["`k_msm_accumulate<G1>` (the roofline kernel): %.1f %% of the proving kernels' time in this list, %.1f %% in the one-proof list above; "
                "event-timed alone 1.52 ms x 4 launches = 6.1 ms of the 19.4 ms single-proof latency (31 %%) -- the proof overlaps its five MSMs on side streams, so the serialised lists weigh the latency-bound tail kernels more than the live run does." % (100 * acc / lib_tot, ONE_PROOF_SHARE.get("acc_g1", float("nan"))), ""]
    census = os.path.join(P, "r2_op_census_2p20.md")
    if os.path.exists(census):
        body = open(census).read().split("\n", 1)[1].strip("\n")
        out += ["## Operation census: the integer-multiply work of one proof, counted (`r2_op_census_2p20.json`, `tools/census_table.py`)", "", body, ""]
    extra = os.path.join(P, "NOTES_r2.md")
    if os.path.exists(extra):
        out += [open(extra).read()]
    open(os.path.join(P, "README.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
