"""Field / curve headers (HOST path) against the big-int oracle.

The same templates (fp2.cuh, ec.cuh) are what the CUDA kernels instantiate; only the
lowest layer (Fp add/sub/mul) differs between host (portable C) and device (PTX carry
chains), and that layer is re-checked on the GPU in tests/test_gpu_field.py.
"""
import ctypes
import random

from oracle import bls12381 as BLS
from oracle import bn128 as O

Q, RR = O.Q, O.R_ORDER
MNT_Q = 0x01C4C62D92C41110229022EEE2CDADB7F997505B8FAFED5EB7E8F96C97D87307FDB925E8A0ED8D99D124D9A15AF79DB26C5C28C859A99B3EEBCA9429212636B9DFF97634993AA4D6C381BC3F0057974EA099170FA13A4FD90776E240000001


def b(x, n=32):
    return x.to_bytes(n, "little")


def i(bs):
    return int.from_bytes(bs, "little")


def edge(p):
    # the reference's edge grid: test/f1.js:299-344
    return [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p - 1) // 2 + 1, (p - 1) // 2 - 1, (1 << 32) - 1, 1 << 32]


def test_fp_ops(host_math):
    rnd = random.Random(1)
    for field, p, n in ((0, Q, 32), (1, RR, 32), (2, MNT_Q, 96), (3, BLS.Q, 48), (4, BLS.R_ORDER, 32)):
        Rm = 1 << (8 * n)
        Rinv = pow(Rm, -1, p)
        vals = edge(p) + [rnd.randrange(p) for _ in range(40)]
        out = ctypes.create_string_buffer(n)
        for x in vals:
            for y in vals[:12] + [rnd.randrange(p) for _ in range(4)]:
                for op, want in ((0, (x + y) % p), (1, (x - y) % p), (2, x * y * Rinv % p)):
                    host_math.h_fp_op(field, op, b(x, n), b(y, n), out)
                    assert i(out.raw) == want, (field, op, x, y)
            host_math.h_fp_op(field, 4, b(x, n), b(0, n), out)
            assert i(out.raw) == (-x) % p
            host_math.h_fp_op(field, 5, b(x, n), b(0, n), out)
            assert i(out.raw) == x * Rm % p
            host_math.h_fp_op(field, 6, b(x, n), b(0, n), out)
            assert i(out.raw) == x * Rinv % p
        for x in vals[:14]:
            host_math.h_fp_op(field, 3, b(x * Rm % p, n), b(0, n), out)
            want = pow(x, -1, p) * Rm % p if x else 0
            assert i(out.raw) == want


def fq2b(v):
    return b(O.to_mont(v[0])) + b(O.to_mont(v[1]))


def fq2i(bs):
    return (O.from_mont(i(bs[:32])), O.from_mont(i(bs[32:64])))


def test_fq2_ops(host_math):
    rnd = random.Random(2)
    out = ctypes.create_string_buffer(64)
    vals = [(0, 0), (1, 0), (0, 1), (Q - 1, Q - 1)] + [(rnd.randrange(Q), rnd.randrange(Q)) for _ in range(20)]
    for x in vals:
        for y in vals:
            for op, fn in ((0, O.Fq2.add), (1, O.Fq2.sub), (2, O.Fq2.mul)):
                host_math.h_fq2_op(op, fq2b(x), fq2b(y), out)
                assert fq2i(out.raw) == fn(x, y)
        host_math.h_fq2_op(5, fq2b(x), fq2b(x), out)
        assert fq2i(out.raw) == O.Fq2.sqr(x)
        if x != (0, 0):
            host_math.h_fq2_op(3, fq2b(x), fq2b(x), out)
            assert fq2i(out.raw) == O.Fq2.inv(x)


class G:
    """Glue for one group: serialisation between the oracle's plain Jacobian triples and the
    kernels' Montgomery XYZZ / affine byte layouts."""

    def __init__(self, lib, g):
        self.lib, self.g = lib, g
        self.F = O.Fq if g == 1 else O.Fq2
        self.fs = 32 * g
        self.pfx = "h_g%d_" % g

    def fb(self, v):
        return b(O.to_mont(v)) if self.g == 1 else fq2b(v)

    def fi(self, bs):
        return O.from_mont(i(bs)) if self.g == 1 else fq2i(bs)

    def affine_bytes(self, P):
        a = O.pt_affine(self.F, P)
        if O.pt_is_zero(self.F, P):
            return bytes(2 * self.fs)           # x == 0 => infinity (build_multiexp.js:335-349)
        return self.fb(a[0]) + self.fb(a[1])

    def call(self, name, *args, outsize=None):
        out = ctypes.create_string_buffer(outsize or 4 * self.fs)
        getattr(self.lib, self.pfx + name)(*args, out)
        return out.raw

    def xyzz(self, P):
        return self.call("from_affine", self.affine_bytes(P))

    def to_affine(self, X):
        r = self.call("to_affine", X, outsize=3 * self.fs)
        return tuple(self.fi(r[k * self.fs:(k + 1) * self.fs]) for k in range(3))

    def want(self, P):
        return O.pt_affine(self.F, P)


def gen_points(F, gen, rnd, n):
    base = (gen[0], gen[1], F.one)
    return [O.pt_times_scalar(F, base, rnd.randrange(1, O.R_ORDER)) for _ in range(n)]


def _group_checks(host_math, g, gen):
    rnd = random.Random(10 + g)
    H = G(host_math, g)
    F = H.F
    pts = gen_points(F, gen, rnd, 6)
    inf = O.pt_zero(F)
    cases = pts + [inf]
    for P in cases:
        XP = H.xyzz(P)
        assert H.to_affine(XP) == H.want(P)
        assert H.to_affine(H.call("dbl", XP)) == H.want(O.pt_double(F, P))
        for Qp in cases + [P, O.pt_neg(F, P)]:
            XQ = H.xyzz(Qp)
            want = H.want(O.pt_add(F, P, Qp))
            assert H.to_affine(H.call("add", XP, XQ)) == want
            if not O.pt_is_zero(F, Qp):
                assert H.to_affine(H.call("madd", XP, H.affine_bytes(Qp))) == want
    # non-trivial Z on both sides: (P+Q) + (Q+R), (P+Q) + (P+Q), (P+Q) - (P+Q)
    P, Qp, Rp = pts[:3]
    a = H.call("add", H.xyzz(P), H.xyzz(Qp))
    c = H.call("add", H.xyzz(Qp), H.xyzz(Rp))
    assert H.to_affine(H.call("add", a, c)) == H.want(O.pt_add(F, O.pt_add(F, P, Qp), O.pt_add(F, Qp, Rp)))
    a2 = H.call("add", H.xyzz(Qp), H.xyzz(P))
    assert H.to_affine(H.call("add", a, a2)) == H.want(O.pt_double(F, O.pt_add(F, P, Qp)))
    an = H.call("add", H.xyzz(O.pt_neg(F, P)), H.xyzz(O.pt_neg(F, Qp)))
    assert H.to_affine(H.call("add", a, an)) == H.want(inf)
    assert H.to_affine(H.call("madd", a, H.affine_bytes(O.pt_add(F, P, Qp)))) == H.want(O.pt_double(F, O.pt_add(F, P, Qp)))
    assert H.to_affine(H.call("madd", a, H.affine_bytes(O.pt_neg(F, O.pt_add(F, P, Qp))))) == H.want(inf)
    # timesScalar: 32-byte and 64-byte scalars, r*G = infinity (test/bn128.js:84-185)
    for k, ln in ((0, 32), (1, 32), (10, 32), (rnd.getrandbits(256), 32), (rnd.getrandbits(512), 64), (O.R_ORDER, 32)):
        got = H.to_affine(H.call("times_scalar", H.xyzz(P), b(k, ln), ctypes.c_int(ln)))
        assert got == H.want(O.pt_times_scalar(F, P, k)), k
    assert H.to_affine(H.call("times_scalar", H.xyzz(P), b(O.R_ORDER, 32), ctypes.c_int(32))) == H.want(inf)


def test_g1_group_law(host_math):
    _group_checks(host_math, 1, O.G1_GEN)


def test_g2_group_law(host_math):
    _group_checks(host_math, 2, O.G2_GEN)


def test_bls12381_group_law_host(host_math):
    """k*P + Q on BLS12-381 G1 / G2 with the same ec.cuh templates the kernels instantiate, vs the affine oracle."""
    rnd = random.Random(9)
    for F, G, enc, fn, size in ((BLS.Fq, BLS.G1_GEN, BLS.g1_bytes, host_math.h_bls_g1_muladd, 144),
                                (BLS.Fq2, BLS.G2_GEN, BLS.g2_bytes, host_math.h_bls_g2_muladd, 288)):
        dec = BLS.g1_from_affine_std if size == 144 else BLS.g2_from_affine_std
        out = ctypes.create_string_buffer(size)
        for k, qk in ((1, 1), (2, 3), (rnd.randrange(1 << 256), rnd.randrange(BLS.R_ORDER)), (BLS.R_ORDER, 5), (BLS.R_ORDER - 7, 7), (5, 5)):
            Qp = BLS.pt_mul(F, G, qk)
            fn(enc(G), b(k, 32), 32, enc(Qp), out)
            want = BLS.pt_add(F, BLS.pt_mul(F, G, k), Qp)
            vals = [i(out.raw[48 * j:48 * j + 48]) for j in range(size // 48)]
            got_bytes = b"".join(b(BLS.from_mont(v), 48) for v in vals)
            assert dec(got_bytes) == want, (size, k, qk)
