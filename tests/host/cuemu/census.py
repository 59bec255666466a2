"""tests/host/cuemu/census.py -- operation census of one Groth16 proof on the CPU kernel emulator (TEST SCAFFOLDING / profiling aid).

    python tests/host/cuemu/census.py [log2 constraints = 20] [out.json]

Runs the bench's synthetic key (tools/gen_synth_key.py, seed 0xB204, nVars = domainSize = 2^k, nPublic = 58) through the
emulated library -- the product's own kernels executed thread by thread -- and counts the Montgomery products and squarings
each kernel executes for ONE proof with a resident key (window tables built beforehand, not counted).  The proof is checked
against tests/golden/synth_2p<k>.golden.json when that file exists.  Used for profiles/r2_op_census_2p20.json: the
integer-multiply work of a proof counted instead of estimated, next to the measured kernel times.
"""
import ctypes
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools"), HERE):
    sys.path.insert(0, p)


def clean_label(label):
    """'k_msm_digits<C> [C = G1Curve; cudaError_t = int]' -> 'k_msm_digits<G1Curve>' (the name ncu prints, minus launch-bound
    arguments); the accumulate kernels are launched through a function-pointer variable: their curve is read off its type"""
    import re
    expr, _, binds = label.partition(" [")
    binds = binds.rstrip("]")
    if expr == "kern":
        m = re.search(r"Affine<(\w+)>", binds)
        return "k_msm_accumulate<%s>" % (m.group(1) if m else "?")
    for item in binds.split("; "):
        k, _, v = item.partition(" = ")
        if re.fullmatch(r"[A-Z]\w*", k) and re.fullmatch(r"\w+", v):
            expr = re.sub(r"\b%s\b" % k, v, expr)
    return expr


def main():
    import build_emu
    import gen_synth_key
    from wasmsnark_b200 import _lib
    log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    out = sys.argv[2] if len(sys.argv) > 2 else None
    lib = _lib.load_library(path=build_emu.build())
    from wasmsnark_b200 import build_bn128
    raw = ctypes.CDLL(build_emu.OUT)
    raw.cuemu_census_json.restype = ctypes.c_size_t
    raw.cuemu_census_json.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    t0 = time.time()
    pk, wit = gen_synth_key.make(1 << log2, 58, 1 << log2, seed=0xB204)
    rs = bytes(range(1, 65))                       # the bench's (r, s)
    bn = build_bn128(0)
    key = bn.load_proving_key(pk)
    t1 = time.time()
    raw.cuemu_census_reset()
    proof = bn.groth16GenProof(wit, key, r=rs[:32], s=rs[32:])
    t2 = time.time()
    n = raw.cuemu_census_json(None, 0)
    buf = ctypes.create_string_buffer(n + 16)
    raw.cuemu_census_json(buf, n + 16)
    census = {}
    for label, c in json.loads(buf.value.decode()).items():
        dst = census.setdefault(clean_label(label), dict.fromkeys(c, 0))
        for k, v in c.items():
            dst[k] += v
    gold = os.path.join(ROOT, "tests", "golden", "synth_2p%d.golden.json" % log2)
    checked = None
    if os.path.exists(gold):
        g = json.load(open(gold))
        want = g.get("proof") or g.get("proof_rs")
        checked = proof == want
    key.free()
    bn.terminate()
    total = {k: sum(c[k] for c in census.values()) for k in ("mul8", "sqr8", "mul12", "sqr12", "mul24", "sqr24", "launches", "threads")}
    res = {"log2_constraints": log2, "equals_reference_golden_proof": checked, "seconds": {"keygen_and_load": round(t1 - t0, 1), "proof": round(t2 - t1, 1)},
           "note": "mulN = every N-limb Montgomery product executed inside kernels (squarings included), sqrN = the squarings among them; one proof, resident key",
           "total": total, "kernels": census}
    text = json.dumps(res, indent=1)
    if out:
        open(out, "w").write(text + "\n")
    print(text if not out else json.dumps({"total": total, "equals_reference_golden_proof": checked, "seconds": res["seconds"]}))


if __name__ == "__main__":
    main()
