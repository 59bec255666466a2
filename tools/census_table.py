"""tools/census_table.py -- joins the emulator's operation census of one 2^20-constraint proof (profiles/r2_op_census_2p20.json,
tests/host/cuemu/census.py: Montgomery products / squarings COUNTED inside the executed kernels) with the measured launch list
of the same proof on a B200 (profiles/r2_launches_prove_2p20.csv, ncu gpu__time_duration) and prints, per kernel, the time the
products alone would take at the multiplier pipe's ceiling next to the measured time.

    python tools/census_table.py > profiles/r2_op_census_2p20.md

Ceiling: IMAD.WIDE issues at 8 lanes / clk / SM sub-partition -> 148 x 4 x 8 x 1.965e9 = 9.31e12 limb products per
second; a 256-bit Montgomery product is 137 of them, a dedicated squaring 111 (SASS counts).  The census counts THREAD-level
products: kernels whose warps are only partly filled (the warp-cooperative chains use 4 of 32 lanes, k_msm_final one warp)
occupy the pipe for whole warp instructions, so for them the "ideal" column is a lower bound by construction.
"""
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
PIPE = 148 * 4 * 8 * 1.965e9          # IMAD.WIDE lanes per second
IMAD_MUL, IMAD_SQR = 137, 111


def key(name):
    name = name.replace("void ", "").replace("ws::", "")
    m = re.match(r"(\w+)(?:<([\w:]+))?", name)
    return m.group(1) + ("<%s>" % m.group(2) if m.group(2) else "")


def measured():
    rows = list(csv.reader(open(os.path.join(P, "r2_launches_prove_2p20.csv"))))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    L = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[start:] if len(r) > mv]
    idx = [i for i, (k, _) in enumerate(L) if "k_assemble_post" in k]
    if len(idx) >= 2:
        L = L[idx[-2] + 1:idx[-1] + 1]            # the launches of the last proof (as tools/make_profiles_r2.py)
    out = collections.OrderedDict()
    for k, ns in L:
        e = out.setdefault(key(re.sub(r"\(.*", "", k)), [0, 0.0])
        e[0] += 1
        e[1] += ns / 1000.0
    return out


def main():
    c = json.load(open(os.path.join(P, "r2_op_census_2p20.json")))
    meas = measured()
    rows, tot_ideal, tot_meas, tot_prod = [], 0.0, 0.0, 0
    for name, k in c["kernels"].items():
        mul, sq = k["mul8"] - k["sqr8"], k["sqr8"]
        ideal_us = (mul * IMAD_MUL + sq * IMAD_SQR) / PIPE * 1e6
        m = meas.get(key(name))
        rows.append((name, k["launches"], mul, sq, ideal_us, m[1] if m else None, m[0] if m else None))
        tot_prod += mul + sq
        tot_ideal += ideal_us
    for name, (n, us) in meas.items():
        tot_meas += us
    rows.sort(key=lambda r: -(r[5] or 0))
    print("# Operation census of one Groth16 proof, 2^20 constraints (BN128, resident key with window tables)")
    print()
    print("Products and squarings are COUNTED inside the product's own kernels, executed thread by thread on the CPU kernel emulator")
    print("(`tests/host/cuemu/census.py`; the emulated proof equals the reference's golden proof: %s); measured times are the ncu launch" % c["equals_reference_golden_proof"])
    print("list of the same proof on a B200 (`r2_launches_prove_2p20.csv`, serialised, cold cache).  `ideal` = the time the counted limb")
    print("products alone occupy the multiplier pipe (137 IMAD.WIDE per product, 111 per dedicated squaring, 9.31e12 lanes/s); thread-level")
    print("counts, so partly filled warps (cooperative chains, single-warp tails) make `ideal` a lower bound by construction.")
    print()
    print("| kernel | launches | products | squarings | ideal us | measured us | ideal / measured |")
    print("|---|---|---|---|---|---|---|")
    for name, n, mul, sq, ideal, mus, mn in rows:
        if mul + sq == 0 and not mus:
            continue
        print("| `%s` | %d | %s | %s | %.1f | %s | %s |" % (name, n, "{:,}".format(mul), "{:,}".format(sq), ideal, "%.1f" % mus if mus else "-",
                                                        "%.0f %%" % (100 * ideal / mus) if mus and ideal else "-"))
    print("| **all kernels of the proof** | %d | %s | | %.0f | %.0f | %.0f %% |" % (c["total"]["launches"], "{:,}".format(tot_prod), tot_ideal, tot_meas, 100 * tot_ideal / tot_meas))
    print()
    print("Total: %.3e Montgomery products per proof = %.2f ms of multiplier-pipe time; measured: %.1f ms serialised under ncu, 17.4 ms per proof"
          % (tot_prod, tot_ideal / 1000, tot_meas / 1000))
    print("with three proofs in flight (`r2_bench_n1.json`) -> the pipelined prover runs at %.0f %% of the integer-multiply ceiling." % (100 * tot_ideal / 1000 / 17.41))


if __name__ == "__main__":
    main()
