"""Pins the oracle (oracle/bn128.py) before anything is checked against it:
  * golden vectors produced by the reference's own code (tests/golden/, see make_golden.py),
  * the reference's shipped golden proof example/bn128/proof_good0.json (r = s = 0),
  * the reference's own WASM image compiled natively (oracle/_ref), when present.
CPU only."""
import json
import os
import random

import pytest

from oracle import bn128 as O
from oracle import ref_wasm

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
EXAMPLE = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "example_bn128")


def fr_list(bs):
    return [O.from_mont(int.from_bytes(bs[i:i + 32], "little"), O.R_ORDER) for i in range(0, len(bs), 32)]


def test_fft_golden_vectors():
    for v in json.load(open(os.path.join(GOLD, "fft_vectors.json"))):
        x = fr_list(bytes.fromhex(v["in"]))
        want = fr_list(bytes.fromhex(v["out"]))
        got = O.fft(x, v["odd"]) if v["fn"] == "fft_fft" else O.ifft(x, v["odd"])
        assert got == want, (v["fn"], v["n"], v["odd"])


def test_fft_identities():
    """test/fft.js:16-121: ifft(fft(x)) = x; even-FFT (+) odd-FFT of size N = FFT of size 2N of the zero-padded input."""
    rnd = random.Random(3)
    for n in (4, 64):
        x = [rnd.randrange(O.R_ORDER) for _ in range(n)]
        assert O.ifft(O.fft(x)) == x
        big = O.fft(x + [0] * n)
        assert big[0::2] == O.fft(x, 0) and big[1::2] == O.fft(x, 1)


def test_montgomery_limb_restatement():
    """f1m_mul restated limb by limb (src/build_f1m.js:235-436) == definitional product; test/f1.js:355-372 closed form."""
    rnd = random.Random(4)
    for p in (O.Q, O.R_ORDER):
        for x, y in [(0, 0), (1, 1), (p - 1, p - 1), (p - 1, 2)] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(200)]:
            assert O.f1m_mul_limbs(x, y, p) == O.mont_mul(x, y, p)
    assert O.to_mont(11, O.R_ORDER) == 11 * (1 << 256) % O.R_ORDER


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not built")
def test_raw_witness_values_reference_vs_oracle():
    """A witness value need not be below r: frm_toMontgomeryN takes the signals as they come (src/bn128.js:134) and the
    multiexps take them as raw 256-bit scalars.  (1) the reference's f1m_toMontgomery / frm_toMontgomery return the canonical
    a R mod p for raw a, the top of the range included; (2) the reference's whole prover on a witness of such values equals the
    big-integer restatement -- which is what the GPU proof of that witness is compared with (test_gpu_prover.py, case
    "witness near 2^256")."""
    rnd = random.Random(3)
    top = (1 << 256) - 1
    inst = ref_wasm.RefInstance(pages=100)
    px, pr = inst.alloc(32), inst.alloc(32)
    for fn, p in (("f1m_toMontgomery", O.Q), ("frm_toMontgomery", O.R_ORDER)):
        for v in [top, top - 1, top ^ 0xFFFFFFFF, top ^ (0xFFFFFFFF << 224), p, p + 1, 5 * p + 1, top - p + 1] + [rnd.randrange(4 * p, 1 << 256) for _ in range(300)]:
            inst.put(px, v.to_bytes(32, "little"))
            inst.call(fn, px, pr)
            assert int.from_bytes(inst.get(pr, 32), "little") == (v << 256) % p == O.to_mont(v, p), (fn, hex(v))
    inst.close()
    pk = open(os.path.join(GOLD, "synth_small.pkey.bin"), "rb").read()
    n = len(open(os.path.join(GOLD, "synth_small.witness.bin"), "rb").read()) // 32
    vals = ([top - 1, top ^ 0xFFFFFFFF, top ^ (0xFFFFFFFF << 224), top - O.R_ORDER + 1] + [rnd.randrange(7 << 253, 1 << 256) for _ in range(n)])[:n]
    wit = b"".join(v.to_bytes(32, "little") for v in vals)
    r, s = bytes(range(7, 39)), bytes(range(100, 132))
    ref = ref_wasm.RefBn128(concurrency=2, pages=400)
    try:
        assert ref.groth16GenProof(wit, pk, r=r, s=s) == O.groth16_gen_proof(wit, pk, int.from_bytes(r, "little"), int.from_bytes(s, "little"))
    finally:
        ref.close()


def test_synth_small_calch_and_proofs():
    pk = open(os.path.join(GOLD, "synth_small.pkey.bin"), "rb").read()
    wit = open(os.path.join(GOLD, "synth_small.witness.bin"), "rb").read()
    g = json.load(open(os.path.join(GOLD, "synth_small.golden.json")))
    k = O.parse_pkey(pk)
    sig = O.rd_scalars(wit, k["nVars"])
    h = O.calc_h(sig, pk[k["pPolsA"]:k["pPolsB"]], pk[k["pPolsB"]:k["pPointsA"]], k["nVars"], k["domainSize"])
    assert b"".join(v.to_bytes(32, "little") for v in h).hex() == g["calcH"]
    assert O.groth16_gen_proof(wit, pk, 0, 0) == g["proof_r0s0"]
    r = int.from_bytes(bytes.fromhex(g["r"]), "little")
    s = int.from_bytes(bytes.fromhex(g["s"]), "little")
    assert O.groth16_gen_proof(wit, pk, r, s) == g["proof_rs"]


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not built")
def test_group_ops_against_reference_image():
    """g1m/g2m add, double, timesScalar, affine of the reference image vs the restatement."""
    rnd = random.Random(6)
    inst = ref_wasm.RefInstance(pages=1000)
    for g, F, gen, size in ((1, O.Fq, O.G1_GEN, 96), (2, O.Fq2, O.G2_GEN, 192)):
        pfx = "g%dm" % g
        enc = (lambda P: O.g1_to_bytes_m(P)) if g == 1 else (lambda P: O.g2_to_bytes_m(P))
        pa, pb, pr, pk = inst.alloc(size), inst.alloc(size), inst.alloc(size), inst.alloc(64)

        def dec(bs):
            v = [O.from_mont(int.from_bytes(bs[i:i + 32], "little")) for i in range(0, size, 32)]
            return tuple(v) if g == 1 else ((v[0], v[1]), (v[2], v[3]), (v[4], v[5]))

        base = (gen[0], gen[1], F.one)
        P = O.pt_times_scalar(F, base, rnd.randrange(1, O.R_ORDER))
        Qp = O.pt_times_scalar(F, base, rnd.randrange(1, O.R_ORDER))
        for A, Bp in ((P, Qp), (P, P), (P, O.pt_neg(F, P)), (P, O.pt_zero(F)), (O.pt_zero(F), Qp)):
            inst.put(pa, enc(A))
            inst.put(pb, enc(Bp))
            inst.call(pfx + "_add", pa, pb, pr)
            inst.call(pfx + "_affine", pr, pr)
            assert dec(inst.get(pr, size)) == O.pt_affine(F, O.pt_add(F, A, Bp))
        kk = rnd.getrandbits(256)
        inst.put(pa, enc(P))
        inst.put(pk, kk.to_bytes(32, "little"))
        inst.call(pfx + "_timesScalar", pa, pk, 32, pr)
        inst.call(pfx + "_affine", pr, pr)
        assert dec(inst.get(pr, size)) == O.pt_affine(F, O.pt_times_scalar(F, P, kk))
    inst.close()


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not built")
def test_multiexp_against_reference_image():
    """g1m_multiexp2 / g2m_multiexp (w = 7) incl. x = 0 entries and unreduced scalars."""
    import util
    rnd = random.Random(7)
    inst = ref_wasm.RefInstance(pages=2000)
    for g, F, gen, enc, dec in ((1, O.Fq, O.G1_GEN, util.g1_affine_m, util.g1_result_tuple), (2, O.Fq2, O.G2_GEN, util.g2_affine_m, util.g2_result_tuple)):
        n = 23
        base = (gen[0], gen[1], F.one)
        pts = [O.pt_times_scalar(F, base, rnd.randrange(1, O.R_ORDER)) for _ in range(n)]
        pts[3] = O.pt_zero(F)
        sc = [rnd.getrandbits(256) for _ in range(n)]
        sc[5], sc[6], sc[7] = 0, 1, (1 << 256) - 1
        sb = b"".join(s.to_bytes(32, "little") for s in sc)
        pb = b"".join(enc(P) for P in pts)
        res = inst.cmd_g1_multiexp(sb, pb, n) if g == 1 else inst.cmd_g2_multiexp(sb, pb, n)
        p0 = inst.alloc(len(res))
        inst.put(p0, res)
        inst.call("g%dm_affine" % g, p0, p0)
        assert dec(inst.get(p0, len(res))) == O.pt_affine(F, O.multiexp_naive(F, sc, pts))
    inst.close()


@pytest.mark.slow
@pytest.mark.skipif(not os.path.exists(os.path.join(EXAMPLE, "proving_key.bin")), reason="example key not staged (oracle/build_ref.py)")
def test_example_proof_good0():
    """The restatement reproduces the reference's golden proof (r = s = 0) on its example key."""
    pk = open(os.path.join(EXAMPLE, "proving_key.bin"), "rb").read()
    wit = open(os.path.join(EXAMPLE, "witness.bin"), "rb").read()
    good = json.load(open(os.path.join(EXAMPLE, "proof_good0.json")))
    gold = json.load(open(os.path.join(GOLD, "example_bn128.golden.json")))
    proof = O.groth16_gen_proof(wit, pk, 0, 0)
    for k in ("pi_a", "pi_b", "pi_c"):
        assert proof[k] == good[k] == gold[k]


@pytest.mark.skipif(not (ref_wasm.available() and os.path.exists(os.path.join(EXAMPLE, "verification_key.json"))), reason="oracle/_ref not staged")
def test_reference_verifier_on_shipped_proofs():
    """The reference's own groth16Verify (src/bn128.js:722-791, pairing code of its WASM image) accepts its shipped
    proofs -- proof_good.json has unknown random r, s -- and rejects a tampered proof / wrong public input."""
    vk = json.load(open(os.path.join(EXAMPLE, "verification_key.json")))
    pub = json.load(open(os.path.join(EXAMPLE, "public.json")))
    for f in ("proof_good.json", "proof_good0.json"):
        assert ref_wasm.groth16_verify(vk, pub, json.load(open(os.path.join(EXAMPLE, f))))
    bad = json.load(open(os.path.join(EXAMPLE, "proof_good.json")))
    bad["pi_c"][0] = str(int(bad["pi_c"][0]) + 1)
    assert not ref_wasm.groth16_verify(vk, pub, bad)
    pub2 = list(pub)
    pub2[3] = str(int(pub2[3]) + 1)
    assert not ref_wasm.groth16_verify(vk, pub2, json.load(open(os.path.join(EXAMPLE, "proof_good.json"))))


@pytest.mark.parametrize("k", [10, 12])
def test_big_integer_prover_equals_the_reference_on_the_bench_key_family(k):
    """oracle/bn128.groth16_gen_proof (the big-integer restatement) against the proof the reference's own code computed
    (tests/golden/make_golden_2p20.py <k>) for the bench's synthetic key family at 2^k constraints with (r, s) != 0 -- the
    same generator, seed and (r, s) as the 2^20 headline key, whose golden proof pins the GPU prover
    (tests/golden/synth_2p20.golden.json).  Also pins the key generator itself: its bytes hash to the stored sha256."""
    import hashlib
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import gen_synth_key
    g = json.load(open(os.path.join(root, "tests", "golden", "synth_2p%d.golden.json" % k)))
    pk, wit = gen_synth_key.make(1 << k, g["n_public"], 1 << k, seed=g["seed"])
    assert hashlib.sha256(pk).hexdigest() == g["pkey_sha256"] and hashlib.sha256(wit).hexdigest() == g["witness_sha256"]
    r, s = (int.from_bytes(bytes.fromhex(g[x]), "little") for x in ("r", "s"))
    assert O.groth16_gen_proof(wit, pk, r, s) == g["proof"]
