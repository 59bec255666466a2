"""Pins oracle/bls12381.py on the reference's own constants and self-tests (no reference image exists for this curve):
generators on their curves, r*G = O (test/bls12381.js:330-338, :381-389), Montgomery helpers, fft <-> ifft."""
import random

from oracle import bls12381 as B


def test_constants_and_generators():
    assert B.Q.bit_length() == 381 and B.R_ORDER.bit_length() == 255
    assert (B.Q - 1) // 64 // 2 ** 0 and (B.Q.bit_length() + 63) // 64 == 6       # n64 = 6 (build_bls12381.js:23)
    assert B.on_curve(B.Fq, B.G1_GEN) and B.on_curve(B.Fq2, B.G2_GEN)
    assert B.pt_mul(B.Fq, B.G1_GEN, B.R_ORDER) is None                           # "G1n == 0"
    assert B.pt_mul(B.Fq2, B.G2_GEN, B.R_ORDER) is None                          # "G2n == 0"
    assert (B.R_ORDER - 1) % (1 << 32) == 0 and (B.R_ORDER - 1) % (1 << 33) != 0
    assert pow(B.NQR, B.R_ORDER >> 1, B.R_ORDER) == B.R_ORDER - 1
    w = B.roots()
    assert w[0] == 1 and w[1] == B.R_ORDER - 1 and pow(w[32], 1 << 31, B.R_ORDER) == B.R_ORDER - 1


def test_group_law_and_multiexp():
    rnd = random.Random(1)
    for F, G in ((B.Fq, B.G1_GEN), (B.Fq2, B.G2_GEN)):
        a, b = rnd.randrange(B.R_ORDER), rnd.randrange(B.R_ORDER)
        P, S = B.pt_mul(F, G, a), B.pt_mul(F, G, b)
        assert B.on_curve(F, P) and B.pt_add(F, P, S) == B.pt_mul(F, G, (a + b) % B.R_ORDER)
        assert B.pt_add(F, P, B.pt_neg(F, P)) is None and B.pt_add(F, P, P) == B.pt_mul(F, G, 2 * a % B.R_ORDER)
        ks = [rnd.randrange(1 << 256) for _ in range(5)]                          # raw 256-bit scalars, not reduced
        pts = [B.pt_mul(F, G, rnd.randrange(B.R_ORDER)) for _ in range(5)]
        assert B.multiexp(F, ks, pts) == B.multiexp(F, [k % B.R_ORDER for k in ks], pts)


def test_montgomery_and_fft():
    rnd = random.Random(2)
    x, y = rnd.randrange(B.Q), rnd.randrange(B.Q)
    assert B.from_mont(B.mont_mul(B.to_mont(x), B.to_mont(y))) == x * y % B.Q
    u, v = rnd.randrange(B.R_ORDER), rnd.randrange(B.R_ORDER)
    assert B.from_mont(B.mont_mul(B.to_mont(u, B.R_ORDER), B.to_mont(v, B.R_ORDER), B.R_ORDER), B.R_ORDER) == u * v % B.R_ORDER
    for n in (1, 2, 8, 64):
        a = [rnd.randrange(B.R_ORDER) for _ in range(n)]
        assert B.ifft(B.fft(a)) == a
        w = B._ROOTS[n.bit_length() - 1]
        assert B.fft(a) == [sum(a[j] * pow(w, j * k, B.R_ORDER) for j in range(n)) % B.R_ORDER for k in range(n)]
        w2 = B._ROOTS[n.bit_length()]
        assert B.fft(a, 1) == [sum(a[j] * pow(w2, j * (2 * k + 1), B.R_ORDER) for j in range(n)) % B.R_ORDER for k in range(n)]
