"""Turns the artefacts a GPU run left in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/make_profiles.py r1"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def launch_table(path, last_marker=None):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    L = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[start:] if len(r) > mv]
    if last_marker:
        idx = [i for i, (k, _) in enumerate(L) if last_marker in k]
        L = L[idx[-1]:]
    a = collections.OrderedDict()
    for k, v in L:
        k = re.sub(r"\(.*", "", k).replace("void ", "").replace("ws::", "")
        a.setdefault(k, [0, 0.0])
        a[k][0] += 1
        a[k][1] += v
    tot = sum(v for _, v in a.values())
    lines = ["| `%s` | %d | %.1f | %.1f %% |" % (k[:60], c, v / 1000.0, 100 * v / tot) for k, (c, v) in sorted(a.items(), key=lambda kv: -kv[1][1])]
    return lines, tot / 1e6, a


def more(b, tag):
    """the other objects of the bench line, the scaling runs and the bench command's own launch list"""
    g = b["groth16"]
    out = ["", "## Other lines of the same bench run", "", "| item | value |", "|---|---|"]
    ex, dr = g.get("example_key"), g.get("dropin_call")
    if ex:
        out.append("| BASELINE configs[3]: `example/bn128` key + witness (nVars 66,232, domain 2^16), r = s = 0 | %.2f ms per proof (%.0f proofs/s), e2e %.2f ms; bit-exact vs `proof_good0.json`: %s; key load %.3f s |"
                   % (ex["ms_per_proof_single_stream"], ex["proofs_per_s"], ex["e2e_ms_per_proof"], ex["bit_exact_vs_proof_good0"], ex["key_load_s"]))
    if dr:
        out.append("| the stateless drop-in call `ws_bn128_groth16_prove(witness, pkey)` at 2^20 (host buffers, key fingerprinted every call) | first call %.2f s, then %.1f ms per call; same proof as the resident path: %s |"
                   % (dr["first_call_s"], dr["repeat_call_ms"], dr["proof_matches_resident_path"]))
    if b.get("mnt6753"):
        out.append("| BASELINE configs[4]: MNT6753 G1 MSM 2^18 (wide-limb correctness path) | %.1f ms (%.2f M pairs/s); closed form: %s |"
                   % (b["mnt6753"]["ms_per_msm"], b["mnt6753"]["pairs_per_s"] / 1e6, b["mnt6753"]["matches_closed_form"]))
    if b.get("bls12381"):
        out.append("| BLS12-381 G1 MSM 2^20 (SURVEY 8f rank 4, parity unpinned) | %.1f ms (%.1f M pairs/s); closed form: %s |"
                   % (b["bls12381"]["ms_per_msm"], b["bls12381"]["pairs_per_s"] / 1e6, b["bls12381"]["matches_closed_form"]))
    if b.get("verifier"):
        v = b["verifier"]
        out.append("| host-side `groth16Verify` on the example proof (SURVEY 8f rank 3), one core | %.1f ms per verification (reference image compiled natively: %s ms); accepts good / rejects tampered: %s |"
                   % (v["ms_per_verify"], "%.1f" % v["reference_ms_per_verify"] if v.get("reference_ms_per_verify") else "n/a", v["accepts_good_rejects_tampered"]))
    rows = []
    for n_, f in ((1, "%s_bench_n1.json"), (2, "%s_bench_n2.json"), (4, "%s_bench_n4.json"), (8, "%s_bench_n8.json")):
        fp = os.path.join(P, f % tag)
        if os.path.exists(fp):
            d = json.load(open(fp))
            gg = d.get("groth16") or {}
            rows.append("| %d | %.1f | %.2f | %.1f | %s | %s |" % (n_, d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6,
                                                                  "%.1f" % gg["value"] if gg else "-", "%.2f" % gg["latency"]["ms_per_proof_single_stream"] if gg else "-"))
    if rows:
        out += ["", "## Scaling runs (torchrun, one rank per GPU; MSM line = weak scaling, Groth16 = one fixed 2^20 proof sharded)", "",
                "| GPUs | MSM M pairs/s | ms per step | e2e M pairs/s | Groth16 proofs/s | proof latency ms |", "|---|---|---|---|---|---|"] + rows
    rf = b.get("roofline", {})
    if rf.get("traffic") and os.path.exists(os.path.join(P, "%s_ncu_accumulate_g1_2p20.txt" % tag)):
        out += ["", "Note: `roofline.traffic` in `%s_bench_n1.json` (%d B) is the constant bench.py carried when that line was recorded; the ncu capture"
                " committed here was taken afterwards on the same code and reads the value in the table above (bench.py now carries it)." % (tag, rf["traffic"])]
    lb = os.path.join(P, "%s_launches_bench_msm.csv" % tag)
    if os.path.exists(lb):
        t, tot, a = launch_table(lb)
        lib = [ln for ln in t if "k_" in ln.split("|")[1]]
        out += ["", "## Launch list of the bench command itself (`%s_launches_bench_msm.csv`)" % tag, "",
                "`ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --no-prove --no-ntt --no-bls --no-cpu-baseline`",
                "(the library's kernels; shares are of all captured launches):", "", "| kernel | launches | us | share |", "|---|---|---|---|"] + lib[:12]
    return "\n".join(out) + "\n"


def main():
    os.makedirs(P, exist_ok=True)
    for src, dst in (("acc_r1.ncu-rep", "%s_ncu_accumulate_g1_2p20.txt"), ("ntt_r1.ncu-rep", "%s_ncu_ntt_pass_2p22.txt")):
        if os.path.exists(os.path.join(G, src)):
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, src)], capture_output=True, text=True).stdout
            open(os.path.join(P, dst % tag), "w").write(out)
    for src, dst in (("launches_msm_r1.csv", "%s_launches_msm_2p20.csv"), ("launches_prove2p20_r1.csv", "%s_launches_prove_2p20.csv"), ("bench_r1_full.json", "%s_bench_n1.json"), ("launches_bench_r1.csv", "%s_launches_bench_msm.csv")):
        if os.path.exists(os.path.join(G, src)):
            open(os.path.join(P, dst % tag), "w").write(open(os.path.join(G, src)).read())
    b = json.load(open(os.path.join(P, "%s_bench_n1.json" % tag)))
    m, mt, ma = launch_table(os.path.join(P, "%s_launches_msm_2p20.csv" % tag), "k_msm_digits")
    p, pt, _ = launch_table(os.path.join(P, "%s_launches_prove_2p20.csv" % tag), "k_assemble_pre")
    acc = [v for k, (c, v) in ma.items() if "accumulate" in k][0] / 1000.0
    ncu = open(os.path.join(P, "%s_ncu_accumulate_g1_2p20.txt" % tag)).read()
    get = lambda key: re.search(re.escape(key) + r"\s+(?:[A-Za-z%/]\S*\s+)?([0-9.]+)\s*$", ncu, re.M).group(1)
    traffic = (float(get("dram__bytes_read.sum ")) + float(get("dram__bytes_write.sum "))) * 1e6
    g = b["groth16"]
    s = """# profiles -- round %s (one B200, sm_100a; clocks %s)

Files
* `%s_bench_n1.json` -- `python bench.py --steps 10 --warmup 3` (N = 1), the line the numbers below come from.
* `%s_ncu_accumulate_g1_2p20.txt` -- `ncu --set full --clock-control none -k regex:k_msm_accumulate` on `tools/run_msm_once.py 20 3` (summary of the raw page, tools/ncu_summary.py).
* `%s_ncu_ntt_pass_2p22.txt` -- same for two `k_ntt_pass64` launches (radix-64 passes) of a 2^22 transform, `tools/run_ntt_once.py 22 1`.
* `%s_launches_msm_2p20.csv`, `%s_launches_prove_2p20.csv` -- `ncu --metrics gpu__time_duration.sum --clock-control none` launch lists (cold-cache, serialised: compare SHARES, not absolutes).
* `r1_sass_fp_mul.txt` -- SASS opcode histogram of the BN128 Montgomery product kernel (`cuobjdump -sass`): 137 IMAD-class instructions, one `IMAD.WIDE.U32(.X)` per limb product; `r1_sass_fp_sqr.txt` -- the same for the dedicated squaring: 111.
* `r1_bench_n2_prove2p18.json`, `r1_bench_n4.json`, `r1_bench_n8.json` -- N = 2 / 4 / 8 runs under torchrun (NCCL) from earlier in the round; `r1_bench_n1_reference_arm.json` -- `bench.py --impl reference`.

## Headline (BASELINE configs[1], G1 MSM n = 2^20, scalars uniform in [0, r))

| quantity | value |
|---|---|
| value (%s, device-resident inputs) | %.1f M pairs/s (%.2f ms per MSM) |
| single-stream latency | %.2f ms per MSM (%.1f M pairs/s) |
| e2e through the host-buffer C ABI (H2D 96 MiB + D2H inside, one host thread per context) | %.1f M pairs/s (%.2f ms per MSM) |
| CPU baseline: the reference's `g1m_multiexp2` (its WASM image compiled natively), %d threads | %.0f pairs/s (single thread %.0f pairs/s) |
| dominant kernel `k_msm_accumulate<G1>` | %.2f ms per launch; 100,663,296 algorithmic B => %.1f GB/s = %.2f %% of the measured 6567.1 GB/s |
| ncu, same kernel | `sm__pipe_fmaheavy_cycles_active` %s %% (the IMAD.WIDE pipe -- the binding resource), SIMT efficiency %s/32, DRAM read+write %.0f MB per launch (%.1fx algorithmic: the 64 MiB point array is gathered once per window), DRAM throughput %s %% of peak |
| Groth16, synthetic 2^20-constraint key, resident key + window tables, 2 proofs in flight | %.1f proofs/s (%.1f ms); single-stream latency %.1f ms; e2e with host witness %.1f proofs/s |
| NTT 2^22 forward + inverse | %.2f ms => %.0f GB/s = %.1f %% of HBM peak (integer-multiply bound) |

Stage stamps of one 2^20 MSM (CUDA events on the launch stream, ms): %s

## Launch list shares -- one G1 MSM, n = 2^20 (%.2f ms serialised under ncu)

| kernel | launches | us | share |
|---|---|---|---|
%s

`k_msm_accumulate`: %.0f %% of the launch list, %.0f %% of the event-timed single-stream step.

## Launch list shares -- one Groth16 proof, 2^20 constraints, window tables (%.1f ms serialised; the real run overlaps the side streams: %.1f ms)

| kernel | launches | us | share |
|---|---|---|---|
%s
%s""" % (tag[1:], json.dumps(b.get("clocks")), tag, tag, tag, tag, tag,
       b["config"]["in_flight"].split(" (")[0], b["value"] / 1e6, b["ms_per_step"], b["latency"]["ms_per_msm_single_stream"], b["latency"]["pairs_per_s"] / 1e6,
       b["e2e"]["value"] / 1e6, b["e2e"]["ms_per_step"], b["cpu_baseline"]["cores"], b["cpu_baseline"]["value"], b["cpu_baseline"]["single_thread_value"],
       b["roofline"]["kernel_ms"], b["roofline"]["achieved"], 100 * b["roofline"]["frac"],
       get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"), get("smsp__thread_inst_executed_per_inst_executed.ratio"),
       traffic / 1e6, traffic / 100663296.0, get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
       g["value"], g["ms_per_proof"], g["latency"]["ms_per_proof_single_stream"], g["e2e"]["value"],
       b["ntt"]["ms_per_round_trip"], b["ntt"]["roofline"]["achieved"], 100 * b["ntt"]["roofline"]["frac"],
       ", ".join("%s %.3f" % kv for kv in b["stages_ms"].items()), mt, "\n".join(m),
       100 * acc / (mt * 1000.0), 100 * b["stages_ms"]["accumulate"] / b["latency"]["ms_per_msm_single_stream"],
       pt, g["latency"]["ms_per_proof_single_stream"], "\n".join(p), more(b, tag))
    open(os.path.join(P, "README.md"), "w").write(s)
    print("traffic bytes per launch:", int(traffic))


if __name__ == "__main__":
    main()
