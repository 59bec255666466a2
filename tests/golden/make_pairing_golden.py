#!/usr/bin/env python3
"""Golden fixtures for the verifier / pairing, from the reference's OWN code (its bn128 WASM image, translated by
oracle/wasm2c.py and driven as src/bn128.js:722-786 drives it).  Needs /root/reference and oracle/_ref built.

  pairing_vectors.json          (a G1, b G2, bn128_pairing(aG1, bG2)) for a few scalars: plain-form coordinates and the
                                12 plain Fq coefficients in ftm order
  example_bn128.verify.json     example/bn128 verification key, public inputs and proofs, with the verdict the
                                reference's groth16Verify returns for the originals and for tampered variants
"""
import copy
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_wasm  # noqa: E402

REF = "/root/reference/example/bn128"
R_ORDER = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def ints(b):
    return [str(int.from_bytes(b[i:i + 32], "little")) for i in range(0, len(b), 32)]


def main():
    m = ref_wasm.RefInstance(pages=1000)
    c = m.call
    rnd = random.Random(0x9A12)
    vec = []
    for a, b in [(1, 1), (2, 3), (rnd.randrange(R_ORDER), rnd.randrange(R_ORDER)), (R_ORDER - 1, 5), (rnd.randrange(R_ORDER), 1)]:
        old = m.heap
        ps, pP, pQ, pE = m.alloc(32), m.alloc(96), m.alloc(192), m.alloc(384)
        m.put(ps, a.to_bytes(32, "little"))
        c("g1m_timesScalar", m.const("pG1gen"), ps, 32, pP)
        c("g1m_affine", pP, pP)
        m.put(ps, b.to_bytes(32, "little"))
        c("g2m_timesScalar", m.const("pG2gen"), ps, 32, pQ)
        c("g2m_affine", pQ, pQ)
        c("bn128_pairing", pP, pQ, pE)
        c("ftm_fromMontgomery", pE, pE)
        c("g1m_fromMontgomery", pP, pP)
        c("g2m_fromMontgomery", pQ, pQ)
        P, Q, E = ints(m.get(pP, 64)), ints(m.get(pQ, 128)), ints(m.get(pE, 384))
        vec.append({"a": str(a), "b": str(b), "P": P, "Q": [Q[0:2], Q[2:4]], "e": E})
        m.heap = old
    json.dump(vec, open(os.path.join(HERE, "pairing_vectors.json"), "w"), indent=1)

    vk = json.load(open(os.path.join(REF, "verification_key.json")))
    vk.pop("vk_alfabeta_12", None)             # unused by groth16Verify's pairingEq4 path (src/bn128.js:781)
    pub = json.load(open(os.path.join(REF, "public.json")))
    cases = []

    def case(name, vk_, pub_, proof_):
        cases.append({"name": name, "public": pub_, "proof": proof_, "vk_patch": None if vk_ is vk else vk_,
                      "valid": bool(ref_wasm.groth16_verify(vk_, pub_, proof_, inst=m))})
    for f in ("proof_good.json", "proof_good0.json"):
        pr = json.load(open(os.path.join(REF, f)))
        pr = {k: pr[k] for k in ("pi_a", "pi_b", "pi_c")}
        case(f, vk, pub, pr)
    good = cases[0]["proof"]
    p2 = list(pub); p2[0] = str((int(p2[0]) + 1) % R_ORDER)
    case("public[0] + 1", vk, p2, good)
    p3 = list(pub); p3[3] = str(R_ORDER)       # input >= r  => false (src/bn128.js:772)
    case("public[3] = r", vk, p3, good)
    bad = copy.deepcopy(good); bad["pi_a"], other = cases[1]["proof"]["pi_a"], None
    case("pi_a of the other proof", vk, pub, bad)
    bad = copy.deepcopy(good); bad["pi_c"] = cases[1]["proof"]["pi_c"]
    case("pi_c of the other proof", vk, pub, bad)
    vk2 = copy.deepcopy(vk); vk2["vk_gamma_2"], vk2["vk_delta_2"] = vk["vk_delta_2"], vk["vk_gamma_2"]
    case("gamma2 <-> delta2", vk2, pub, good)
    assert [x["valid"] for x in cases] == [True, True, False, False, False, False, False], [x["valid"] for x in cases]
    json.dump({"verification_key": vk, "cases": cases}, open(os.path.join(HERE, "example_bn128.verify.json"), "w"), indent=1)
    print("pairing / verifier fixtures written")


if __name__ == "__main__":
    main()
