"""Groth16 verifier + BN128 pairing (host code of the C ABI).  No GPU needed.

Golden data: tests/golden/pairing_vectors.json and example_bn128.verify.json, produced by the reference's own image
(tests/golden/make_pairing_golden.py): bn128_pairing values and groth16Verify verdicts (src/bn128.js:722-786) for
the shipped example proofs and tampered variants.  test/bn128.js:187-230 (pairing bilinearity) is mirrored below."""
import copy
import json
import os
import random

import pytest

from wasmsnark_b200 import groth16Verify, pairing
from wasmsnark_b200._lib import WsError

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def test_pairing_matches_reference_image_bit_for_bit():
    vec = json.load(open(os.path.join(GOLD, "pairing_vectors.json")))
    assert len(vec) >= 5
    for v in vec:
        assert [str(x) for x in pairing(v["P"], v["Q"])] == v["e"]


def f12_one():
    return [1] + [0] * 11


def test_pairing_bilinear_and_degenerate():
    """e(O, Q) = e(P, O) = 1 and e(-P, Q) = e(P, Q)^-1 (the conjugate, in the cyclotomic subgroup); products of
    pairings are exercised by the verifier equation below (test/bn128.js:187-230 checks e(aG1, bG2) = e(G1, abG2))."""
    vec = json.load(open(os.path.join(GOLD, "pairing_vectors.json")))
    by = {(int(v["a"]), int(v["b"])): v for v in vec}
    g = by[(1, 1)]
    assert pairing([0, 0], g["Q"]) == f12_one()
    assert pairing(g["P"], [[0, 0], [0, 0]]) == f12_one()
    # e((r-1) G, 5 H) = e(-G, 5H) = conj(e(G, 5H)): the negation of P flips the sign of the c1 half
    v = by[(R - 1, 5)]
    neg_p = [v["P"][0], str((Q - int(v["P"][1])) % Q)]
    e1, e2 = pairing(v["P"], v["Q"]), pairing(neg_p, v["Q"])
    assert e1[:6] == e2[:6] and all((a + b) % Q == 0 for a, b in zip(e1[6:], e2[6:]))


def test_pairing_of_twist_points_outside_the_subgroup_equals_the_reference():
    """The reference never checks that a G2 input lies in the order-r subgroup and runs its projective Miller loop on any
    point of the twist curve; the affine walk with a shared batch inversion here produces the same bytes for such points
    (golden values from the reference image, tests/golden/make_pairing_nonsubgroup_golden.py), and a verification whose
    pi_b is such a point is rejected like any other wrong proof -- the trusted pairs of the product are not disturbed."""
    vec = json.load(open(os.path.join(GOLD, "pairing_vectors_nonsubgroup.json")))
    assert len(vec) >= 4
    for v in vec:
        assert [str(x) for x in pairing(v["P"], v["Q"])] == v["e"]
    g = json.load(open(os.path.join(GOLD, "example_bn128.verify.json")))
    vk, good = g["verification_key"], g["cases"][0]
    assert groth16Verify(vk, good["public"], good["proof"])
    bad = copy.deepcopy(good["proof"])
    bad["pi_b"] = [vec[0]["Q"][0], vec[0]["Q"][1], ["1", "0"]]
    assert not groth16Verify(vk, good["public"], bad)
    assert groth16Verify(vk, good["public"], good["proof"])          # and the next good proof still verifies


def test_pairing_rejects_points_off_curve():
    vec = json.load(open(os.path.join(GOLD, "pairing_vectors.json")))
    v = vec[1]
    with pytest.raises(WsError):
        pairing([v["P"][0], str(int(v["P"][1]) + 1)], v["Q"])
    with pytest.raises(WsError):
        pairing(v["P"], [[v["Q"][0][0], v["Q"][0][1]], [v["Q"][1][0], str(int(v["Q"][1][1]) + 1)]])
    with pytest.raises(WsError):
        pairing([str(Q), v["P"][1]], v["Q"])                 # coordinate >= q


def test_verify_same_verdicts_as_reference():
    g = json.load(open(os.path.join(GOLD, "example_bn128.verify.json")))
    vk = g["verification_key"]
    seen = set()
    for c in g["cases"]:
        got = groth16Verify(c["vk_patch"] or vk, c["public"], c["proof"])
        assert got == c["valid"], c["name"]
        seen.add(c["valid"])
    assert seen == {True, False}


def test_proof_json_shape_matches_the_shipped_example():
    from wasmsnark_b200 import proof_json
    g = json.load(open(os.path.join(GOLD, "example_bn128.verify.json")))
    good = g["cases"][0]["proof"]
    back = json.loads(proof_json(good))
    assert list(back.keys()) == ["pi_a", "pi_b", "pi_c", "protocol"] and back["protocol"] == "groth"      # example/bn128/proof_good.json
    assert groth16Verify(g["verification_key"], g["cases"][0]["public"], back)


def test_verify_tamper_sweep():
    """flipping any single proof coordinate to another curve point's, or any one public input, must reject"""
    g = json.load(open(os.path.join(GOLD, "example_bn128.verify.json")))
    vk, good, other = g["verification_key"], g["cases"][0], g["cases"][1]
    assert groth16Verify(vk, good["public"], good["proof"])
    rnd = random.Random(5)
    for i in rnd.sample(range(len(good["public"])), 6):
        pub = list(good["public"])
        pub[i] = str((int(pub[i]) + rnd.randrange(1, R)) % R)
        assert not groth16Verify(vk, pub, good["proof"])
    for k in ("pi_a", "pi_b", "pi_c"):
        bad = copy.deepcopy(good["proof"])
        bad[k] = other["proof"][k]
        assert not groth16Verify(vk, good["public"], bad)
    bad = copy.deepcopy(good["proof"])
    bad["pi_a"] = [bad["pi_a"][0], str(int(bad["pi_a"][1]) ^ 1), "1"]      # off the curve
    assert not groth16Verify(vk, good["public"], bad)
    with pytest.raises(ValueError):
        groth16Verify(vk, good["public"][:-1], good["proof"])


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libbn128_ref.so")), reason="reference image not built")
def test_verify_and_pairing_against_live_reference_image():
    """fresh random cases against the reference image itself (not only the committed vectors)"""
    from oracle import ref_wasm
    m = ref_wasm.RefInstance(pages=1000)
    c = m.call
    rnd = random.Random(77)
    for _ in range(3):
        a, b = rnd.randrange(R), rnd.randrange(R)
        old = m.heap
        ps, pP, pQ, pE = m.alloc(32), m.alloc(96), m.alloc(192), m.alloc(384)
        m.put(ps, a.to_bytes(32, "little")); c("g1m_timesScalar", m.const("pG1gen"), ps, 32, pP); c("g1m_affine", pP, pP)
        m.put(ps, b.to_bytes(32, "little")); c("g2m_timesScalar", m.const("pG2gen"), ps, 32, pQ); c("g2m_affine", pQ, pQ)
        c("bn128_pairing", pP, pQ, pE); c("ftm_fromMontgomery", pE, pE)
        c("g1m_fromMontgomery", pP, pP); c("g2m_fromMontgomery", pQ, pQ)
        ints = lambda bs: [int.from_bytes(bs[i:i + 32], "little") for i in range(0, len(bs), 32)]   # noqa: E731
        P, Qq, E = ints(m.get(pP, 64)), ints(m.get(pQ, 128)), ints(m.get(pE, 384))
        m.heap = old
        assert pairing(P, [Qq[0:2], Qq[2:4]]) == E
    g = json.load(open(os.path.join(GOLD, "example_bn128.verify.json")))
    vk, good = g["verification_key"], g["cases"][0]
    for i in range(4):
        pub = list(good["public"])
        if i % 2:
            pub[i] = str(rnd.randrange(R))
        assert groth16Verify(vk, pub, good["proof"]) == ref_wasm.groth16_verify(vk, pub, good["proof"], inst=m)
    m.close()
