#!/usr/bin/env python3
"""Generates the committed golden fixtures from the reference's OWN code: its WASM image
translated to C (oracle/wasm2c.py -> oracle/_ref/libbn128_ref.so) driven exactly as the
reference's JS host drives it (oracle/ref_wasm.py).  Run in the build container (needs
/root/reference); outputs are small files under tests/golden/.

  fft_vectors.json      fft_fft / fft_ifft (odd = 0, 1) on n = 4, 16, 64 (test/fft.js sizes + odd variants)
  synth_small.*         synthetic key (nVars 300, nPublic 5, domain 256) + witness, with the reference's
                        calcH output and proofs for r = s = 0 and one fixed (r, s)
  proof_good0.json      the reference's shipped golden proof (example/bn128/proof_good0.json), re-derived here
                        from its key + witness with r = s = 0 and asserted equal before being written
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import bn128 as O  # noqa: E402
from oracle import ref_wasm  # noqa: E402
import gen_synth_key  # noqa: E402

REF = "/root/reference"


def main():
    inst = ref_wasm.RefInstance(pages=2000)
    rnd = random.Random(0xF7)
    vec = []
    for n in (1, 2, 4, 16, 64):
        x = [O.to_mont(rnd.randrange(O.R_ORDER), O.R_ORDER) for _ in range(n)]
        xb = b"".join(v.to_bytes(32, "little") for v in x)
        for fn in ("fft_fft", "fft_ifft"):
            for odd in (0, 1):
                if n == 1 and fn == "fft_ifft":
                    continue        # the reference's __finalInverse never terminates for n = 1 (loop runs i = 1 .. n/2 = 0)
                old = inst.heap
                p = inst.put_bin(xb)
                inst.call(fn, p, n, odd)
                out = inst.get(p, 32 * n)
                inst.heap = old
                vec.append({"fn": fn, "n": n, "odd": odd, "in": xb.hex(), "out": out.hex()})
    json.dump(vec, open(os.path.join(HERE, "fft_vectors.json"), "w"))

    pk, wit = gen_synth_key.make(300, 5, 256, seed=0xB204)
    open(os.path.join(HERE, "synth_small.pkey.bin"), "wb").write(pk)
    open(os.path.join(HERE, "synth_small.witness.bin"), "wb").write(wit)
    B = ref_wasm.RefBn128()
    k = O.parse_pkey(pk)
    h = B.calcH(wit, pk[k["pPolsA"]:k["pPolsB"]], pk[k["pPolsB"]:k["pPointsA"]], k["nVars"], k["domainSize"])
    r = bytes(rnd.getrandbits(8) for _ in range(32))
    s = bytes(rnd.getrandbits(8) for _ in range(32))
    gold = {
        "calcH": h.hex(),
        "proof_r0s0": B.groth16GenProof(wit, pk, r=bytes(32), s=bytes(32)),
        "r": r.hex(), "s": s.hex(),
        "proof_rs": B.groth16GenProof(wit, pk, r=r, s=s),
    }
    json.dump(gold, open(os.path.join(HERE, "synth_small.golden.json"), "w"), indent=1)

    wit_e = open(os.path.join(REF, "example/bn128/witness.bin"), "rb").read()
    pk_e = open(os.path.join(REF, "example/bn128/proving_key.bin"), "rb").read()
    proof = B.groth16GenProof(wit_e, pk_e, r=bytes(32), s=bytes(32))
    good = json.load(open(os.path.join(REF, "example/bn128/proof_good0.json")))
    assert all(proof[kk] == good[kk] for kk in ("pi_a", "pi_b", "pi_c")), "reference image does not reproduce proof_good0.json"
    rs_proof = B.groth16GenProof(wit_e, pk_e, r=r, s=s)
    json.dump({"pi_a": good["pi_a"], "pi_b": good["pi_b"], "pi_c": good["pi_c"], "r": r.hex(), "s": s.hex(), "proof_rs": rs_proof},
              open(os.path.join(HERE, "example_bn128.golden.json"), "w"), indent=1)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
