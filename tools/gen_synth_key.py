#!/usr/bin/env python3
"""Synthetic proving_key.bin / witness.bin in the reference's binary layouts
(tools/buildpkey.js:130-183, tools/buildwitness.js:36-69), for sizes the shipped example
(domain 2^16) does not cover -- e.g. the 2^20-constraint scale-up of BASELINE.json.

Not a valid Groth16 key (no trapdoor structure): the prover's arithmetic does not care, and the
oracle / reference produce the proof of the same bytes for comparison.  Point sections:
  S_i = P0 + i*D on BN128 G1 (one batched-inversion chord addition each);
  A = S, B1 = S rotated by 1, C = S rotated by 2, hExps = S rotated by 3,
  B2_i = (x_i * xi^2, y_i * xi^3) with xi in Fq2: a point of the a = 0 curve y^2 = x^3 + 3 xi^6 over
  Fq2 with both components of every coordinate non-zero (the group law never uses b).
A share of the B1/B2/A entries is the point at infinity (x = 0), like a real key.
Polynomial sections: signal 0 (the constant 1) owns a long list, the others 0..3 terms.
Witness: 6.7 % zeros, 3.1 % ones, 10 % below 2^32, rest uniform below r.
"""
import os
import random
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import bn128 as O  # noqa: E402


def _b32(x):
    return x.to_bytes(32, "little")


def make(n_vars, n_public, domain, seed=0xB204, inf_share=0.3, coef_density=(1.5, 2.1)):
    import util
    rnd = random.Random(seed)
    npts = max(n_vars, domain, 4) + 4          # the five fixed points take entries 5..7 (tiny keys included)
    aff, P0, D = util.arith_points(O.Fq, O.G1_GEN, npts, seed ^ 0x55)
    g1 = [_b32(O.to_mont(x)) + _b32(O.to_mont(y)) for x, y in aff]
    xi = (rnd.randrange(1, O.Q), rnd.randrange(1, O.Q))
    xi2 = O.Fq2.sqr(xi)
    xi3 = O.Fq2.mul(xi2, xi)
    g2 = []
    for x, y in aff[:max(n_vars, 7)]:
        X = (xi2[0] * x % O.Q, xi2[1] * x % O.Q)
        Y = (xi3[0] * y % O.Q, xi3[1] * y % O.Q)
        g2.append(b"".join(_b32(O.to_mont(c)) for c in (X[0], X[1], Y[0], Y[1])))
    inf1, inf2 = bytes(64), bytes(128)
    A = [g1[i] if rnd.random() > inf_share / 2 else inf1 for i in range(n_vars)]
    infB = [rnd.random() < inf_share for _ in range(n_vars)]
    B1 = [inf1 if infB[i] else g1[i + 1] for i in range(n_vars)]
    B2 = [inf2 if infB[i] else g2[i] for i in range(n_vars)]
    C = [g1[i + 2] for i in range(n_vars - n_public - 1)]
    H = [g1[i + 3] for i in range(domain)]

    def pols(density, heavy):
        out = []
        for i in range(n_vars):
            if i == 0:
                nc = min(domain, heavy)
                idxs = sorted(rnd.sample(range(domain), nc))
            else:
                nc = min(domain, max(0, int(rnd.expovariate(1.0 / density)) if density > 0 else 0)) if rnd.random() < 0.7 else 0
                idxs = sorted(rnd.sample(range(domain), nc))
            out.append(struct.pack("<I", nc))
            for idx in idxs:
                coef = rnd.choice((1, O.R_ORDER - 1, rnd.randrange(O.R_ORDER)))
                out.append(struct.pack("<I", idx) + _b32(O.to_mont(coef, O.R_ORDER)))
        return b"".join(out)

    polsA = pols(coef_density[0], max(1, domain // 6))
    polsB = pols(coef_density[1], max(1, domain // 3))
    fixed = b"".join(g1[k] for k in (5, 6, 7)) + g2[5] + g2[6]       # alfa1, beta1, delta1, beta2, delta2
    pA = 40 + len(fixed)
    assert pA == 488
    pB = pA + len(polsA)
    pPA = pB + len(polsB)
    pPB1 = pPA + 64 * n_vars
    pPB2 = pPB1 + 64 * n_vars
    pPC = pPB2 + 128 * n_vars
    pPH = pPC + 64 * len(C)
    hdr = struct.pack("<10I", n_vars, n_public, domain, pA, pB, pPA, pPB1, pPB2, pPC, pPH)
    pkey = b"".join([hdr, fixed, polsA, polsB] + A + B1 + B2 + C + H)
    wit = [1]
    for _ in range(n_vars - 1):
        u = rnd.random()
        if u < 0.067:
            wit.append(0)
        elif u < 0.098:
            wit.append(1)
        elif u < 0.198:
            wit.append(rnd.getrandbits(32))
        else:
            wit.append(rnd.randrange(O.R_ORDER))
    witness = b"".join(_b32(w) for w in wit)
    return pkey, witness


if __name__ == "__main__":
    nv, npub, dom = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    pk, w = make(nv, npub, dom)
    open(sys.argv[4], "wb").write(pk)
    open(sys.argv[5], "wb").write(w)
    print("pkey %d bytes, witness %d bytes" % (len(pk), len(w)))
