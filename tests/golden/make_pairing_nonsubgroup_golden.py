#!/usr/bin/env python3
"""Golden pairing values for twist points OUTSIDE the order-r subgroup, from the reference's own image (bn128_pairing,
src/bn128/build_bn128.js:1320-1346).  The reference does not check subgroup membership and computes f_{6z+2,Q}(P) with
its projective Miller loop for any point of the twist curve; the host verifier here walks the twist point in affine
coordinates with a shared batch inversion (could that diverge, or corrupt the other pairs of the product?).
These vectors pin that both give the same bytes.  Needs /root/reference mounted and oracle/_ref built.

    python tests/golden/make_pairing_nonsubgroup_golden.py      -> tests/golden/pairing_vectors_nonsubgroup.json"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import bn128 as O  # noqa: E402
from oracle import ref_wasm  # noqa: E402

Q, F2 = O.Q, O.Fq2


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = F2.mul(r, a)
        a = F2.sqr(a)
        e >>= 1
    return r


def f2_sqrt(a):
    """q = 3 mod 4 (Adj, Rodriguez-Henriquez: square roots in even extension fields, algorithm 9); None if no root."""
    if a == (0, 0):
        return a
    a1 = f2_pow(a, (Q - 3) // 4)
    alpha = F2.mul(F2.sqr(a1), a)
    a0 = F2.mul((alpha[0], (-alpha[1]) % Q), alpha)
    if a0 == (Q - 1, 0):
        return None
    x0 = F2.mul(a1, a)
    if alpha == (Q - 1, 0):
        return F2.mul((0, 1), x0)
    return F2.mul(f2_pow(F2.add((1, 0), alpha), (Q - 1) // 2), x0)


def twist_points(n, seed):
    """random points of y^2 = x^3 + 3/(9 + u) over Fq2: outside the order-r subgroup with overwhelming probability (checked)."""
    b = F2.mul((3, 0), F2.inv((9, 1)))
    rnd = random.Random(seed)
    out = []
    while len(out) < n:
        x = (rnd.randrange(Q), rnd.randrange(Q))
        rhs = F2.add(F2.mul(F2.sqr(x), x), b)
        y = f2_sqrt(rhs)
        if y is None or F2.sqr(y) != rhs:
            continue
        assert not O.pt_is_zero(F2, O.pt_times_scalar(F2, (x, y, (1, 0)), O.R_ORDER)), "landed in the subgroup"
        out.append((x, y))
    return out


def main():
    m = ref_wasm.RefInstance(pages=1000)
    c = m.call
    rnd = random.Random(0x51B)
    vec = []
    for x, y in twist_points(4, 0x7715):
        k = rnd.randrange(1, O.R_ORDER)
        P = O.pt_affine(O.Fq, O.pt_times_scalar(O.Fq, (O.G1_GEN[0], O.G1_GEN[1], 1), k))
        old = m.heap
        pP, pQ, pE = m.alloc(96), m.alloc(192), m.alloc(384)
        m.put(pP, b"".join(v.to_bytes(32, "little") for v in (P[0], P[1], 1)))
        c("g1m_toMontgomery", pP, pP)
        m.put(pQ, b"".join(v.to_bytes(32, "little") for v in (x[0], x[1], y[0], y[1], 1, 0)))
        c("g2m_toMontgomery", pQ, pQ)
        c("bn128_pairing", pP, pQ, pE)
        c("ftm_fromMontgomery", pE, pE)
        e = m.get(pE, 384)
        m.heap = old
        vec.append({"P": [str(P[0]), str(P[1])], "Q": [[str(x[0]), str(x[1])], [str(y[0]), str(y[1])]],
                    "e": [str(int.from_bytes(e[i:i + 32], "little")) for i in range(0, 384, 32)]})
    json.dump(vec, open(os.path.join(HERE, "pairing_vectors_nonsubgroup.json"), "w"), indent=1)
    print("wrote %d vectors" % len(vec))


if __name__ == "__main__":
    main()
