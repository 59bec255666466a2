"""BLS12-381 instantiation on the GPU through the C ABI, against oracle/bls12381.py.
Parity is UNPINNED for this curve (no reference image exists, see the oracle's header): the checks are against the
big-integer restatement and closed forms."""
import random

import pytest

from oracle import bls12381 as B

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bls():
    from wasmsnark_b200.bls12381 import build_bls12381
    b = build_bls12381(0)
    yield b
    b.terminate()


def le(v, n):
    return int(v).to_bytes(n, "little")


def ints(b, n):
    return [int.from_bytes(b[i:i + n], "little") for i in range(0, len(b), n)]


@pytest.mark.parametrize("which", ["fq", "fr"])
def test_field_ops(bls, which):
    from wasmsnark_b200 import _lib
    p, fid, esz = (B.Q, _lib.FIELD_BLS12381_FQ, 48) if which == "fq" else (B.R_ORDER, _lib.FIELD_BLS12381_FR, 32)
    Rm = 1 << (8 * esz)
    rnd = random.Random(3)
    xs = [0, 1, p - 1, p - 2, Rm % p] + [rnd.randrange(p) for _ in range(600)]
    ys = [p - 1, 0, p - 1, 1, Rm % p] + [rnd.randrange(p) for _ in range(600)]
    X, Y = b"".join(le(v, esz) for v in xs), b"".join(le(v, esz) for v in ys)
    rinv = pow(Rm, -1, p)
    assert ints(bls.fp_batch(fid, _lib.OP_MUL, X, Y), esz) == [a * b * rinv % p for a, b in zip(xs, ys)]
    assert ints(bls.fp_batch(fid, _lib.OP_ADD, X, Y), esz) == [(a + b) % p for a, b in zip(xs, ys)]
    assert ints(bls.fp_batch(fid, _lib.OP_SUB, X, Y), esz) == [(a - b) % p for a, b in zip(xs, ys)]
    assert ints(bls.fp_batch(fid, _lib.OP_SQUARE, X), esz) == [a * a * rinv % p for a in xs]
    assert ints(bls.fp_batch(fid, _lib.OP_TOMONT, X), esz) == [a * Rm % p for a in xs]
    assert ints(bls.fp_batch(fid, _lib.OP_FROMMONT, X), esz) == [a * rinv % p for a in xs]
    # f1m_inverse on Montgomery residues: (a R)^-1 R^2 ... = a^-1 R; inverse(0) = 0
    assert ints(bls.fp_batch(fid, _lib.OP_INVERSE, X), esz) == [pow(a * rinv % p, -1, p) * Rm % p if a else 0 for a in xs]


@pytest.mark.parametrize("which", ["fq", "fr"])
def test_to_montgomery_takes_raw_values(bls, which):
    """f1m_toMontgomery on ANY number that fits the element (as for BN128, tests/test_gpu_field.py): the canonical x R mod p,
    the top of the range included -- the device product takes the raw value as its second operand (fp.cuh::to_mont)"""
    from wasmsnark_b200 import _lib
    p, fid, esz = (B.Q, _lib.FIELD_BLS12381_FQ, 48) if which == "fq" else (B.R_ORDER, _lib.FIELD_BLS12381_FR, 32)
    Rm = 1 << (8 * esz)
    top = Rm - 1
    rnd = random.Random(5)
    xs = [top, top - 1, top ^ 0xFFFFFFFF, top ^ (0xFFFFFFFF << (8 * esz - 32)), top - p, top - p + 1, p, p + 1, 2 * p + 1] + \
         [rnd.randrange(top - 2 * p if top > 4 * p else p, Rm) for _ in range(600)] + [rnd.randrange(Rm) for _ in range(200)]
    assert ints(bls.fp_batch(fid, _lib.OP_TOMONT, b"".join(le(v, esz) for v in xs)), esz) == [a * Rm % p for a in xs]


def test_fq2_mul(bls):
    rnd = random.Random(4)
    xs = [(rnd.randrange(B.Q), rnd.randrange(B.Q)) for _ in range(300)] + [(0, 0), (1, 0), (0, 1), (B.Q - 1, B.Q - 1)]
    ys = [(rnd.randrange(B.Q), rnd.randrange(B.Q)) for _ in range(300)] + [(5, 7), (B.Q - 1, 3), (0, 1), (B.Q - 1, B.Q - 1)]
    enc = lambda vs: b"".join(le(B.to_mont(c), 48) for v in vs for c in v)   # noqa: E731
    got = ints(bls.fq2_mul_batch(enc(xs), enc(ys)), 48)
    want = [B.to_mont(c) for a, b in zip(xs, ys) for c in B.Fq2.mul(a, b)]
    assert got == want


def arith_points(F, G, a, d, n):
    """P_i = (a + i d) G by repeated affine addition."""
    P, D = B.pt_mul(F, G, a), B.pt_mul(F, G, d)
    out = []
    for _ in range(n):
        out.append(P)
        P = B.pt_add(F, P, D)
    return out


def msm_case(bls, F, G, n, rnd, special=False):
    g1 = F is B.Fq
    a, d = rnd.randrange(B.R_ORDER), rnd.randrange(B.R_ORDER)
    pts = arith_points(F, G, a, d, n)
    ks = [rnd.randrange(1 << 256) for _ in range(n)]                       # raw, unreduced 256-bit scalars
    coef = [(a + i * d) % B.R_ORDER for i in range(n)]
    if special and n >= 8:
        ks[0], ks[1], ks[2], ks[3] = 0, 1, B.R_ORDER, (1 << 256) - 1
        pts[4], coef[4] = None, 0                                           # x == 0 => infinity
        pts[6], coef[6] = pts[5], coef[5]                                   # repeated point
        ks[7] = ks[5]                                                       # ... with the same scalar (P + P in a bucket)
        pts[8 % n], coef[8 % n] = B.pt_neg(F, pts[5]), (-coef[5]) % B.R_ORDER
    enc = B.g1_bytes if g1 else B.g2_bytes
    sb = b"".join(le(k, 32) for k in ks)
    pb = b"".join(enc(P) for P in pts)
    res = bls.g1_multiexp(sb, pb) if g1 else bls.g2_multiexp(sb, pb)
    aff = (bls.g1_affine if g1 else bls.g2_affine)(res, 1)
    got = (B.g1_from_affine_std if g1 else B.g2_from_affine_std)(aff)
    want = B.pt_mul(F, G, sum(k * c for k, c in zip(ks, coef)) % B.R_ORDER)
    assert got == want
    return ks, pts, got


@pytest.mark.parametrize("n", [0, 1, 2, 33, 1000])
def test_g1_multiexp_small(bls, n):
    rnd = random.Random(100 + n)
    ks, pts, got = msm_case(bls, B.Fq, B.G1_GEN, n, rnd, special=True)
    if 0 < n <= 33:
        assert got == B.multiexp(B.Fq, ks, pts)                             # definitional sum, term by term


@pytest.mark.parametrize("n", [0, 1, 9, 300])
def test_g2_multiexp_small(bls, n):
    rnd = random.Random(200 + n)
    ks, pts, got = msm_case(bls, B.Fq2, B.G2_GEN, n, rnd, special=True)
    if 0 < n <= 9:
        assert got == B.multiexp(B.Fq2, ks, pts)


def test_g1_multiexp_2p16_closed_form(bls):
    msm_case(bls, B.Fq, B.G1_GEN, 1 << 16, random.Random(7))


def test_g2_multiexp_2p12_closed_form(bls):
    msm_case(bls, B.Fq2, B.G2_GEN, 1 << 12, random.Random(8))


def test_affine_of_infinity_and_generator(bls):
    one = le(B.to_mont(1), 48)
    inf = bytes(48) + one + bytes(48)                                        # (0, 1, 0)
    x, y = B.G1_GEN
    # Jacobian (x z^2, y z^3, z): plain arithmetic, then to Montgomery form
    zp = 12345
    jac = le(B.to_mont(x * zp * zp % B.Q), 48) + le(B.to_mont(y * zp ** 3 % B.Q), 48) + le(B.to_mont(zp), 48)
    out = bls.g1_affine(inf + jac, 1)
    assert B.g1_from_affine_std(out[:144]) is None
    assert B.g1_from_affine_std(out[144:]) == B.G1_GEN
    out_m = bls.g1_affine(jac, 0)                                            # stays in Montgomery form
    assert ints(out_m, 48) == [B.to_mont(x), B.to_mont(y), B.to_mont(1)]


@pytest.mark.parametrize("logn", [0, 1, 2, 3, 5, 6, 7, 10, 12])
def test_fft_vs_oracle(bls, logn):
    n = 1 << logn
    rnd = random.Random(300 + logn)
    xs = [rnd.randrange(B.R_ORDER) for _ in range(n)]
    mont = lambda vs: b"".join(le(B.to_mont(v, B.R_ORDER), 32) for v in vs)   # noqa: E731
    for odd in (0, 1):
        assert bls.fft(mont(xs), odd) == mont(B.fft(xs, odd))
        assert bls.ifft(mont(xs), odd) == mont(B.ifft(xs, odd))


def test_fft_round_trip_2p20_and_bad_sizes(bls):
    from wasmsnark_b200._lib import WsError
    import numpy as np
    n = 1 << 20
    rng = np.random.default_rng(5)
    x = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    x[:, 31] &= 0x3F                                                         # < 2^254 < r: canonical residues
    xb = x.tobytes()
    y = bls.fft(xb)
    assert y != xb and bls.ifft(y) == xb
    with pytest.raises(WsError):
        bls.fft(bytes(32 * 3))


def test_reference_group_relations_through_the_msm_abi(bls):
    """The reference's own BLS12-381 group tests hold no vectors, only relations (test/bls12381.js:289-389): 4G by
    additions == 4G by doublings, 10 G by timesScalar == ten additions of G, r G == infinity, for G1 and G2, on the
    generators its module builds (src/bls12381/build_bls12381.js:117-152).  The same relations, evaluated by the GPU
    kernels through the MSM entry points (a sum of k copies of G with scalar 1 is k additions; a single pair is a
    timesScalar).  This is all the reference offers for this curve: parity stays UNPINNED."""
    for g1 in (True, False):
        G = B.G1_GEN if g1 else B.G2_GEN
        gb = B.g1_bytes(G) if g1 else B.g2_bytes(G)
        msm = bls.g1_multiexp if g1 else bls.g2_multiexp
        aff = bls.g1_affine if g1 else bls.g2_affine
        dec = B.g1_from_affine_std if g1 else B.g2_from_affine_std
        one, two, four, ten = (le(v, 32) for v in (1, 2, 4, 10))

        def point(scalars, npts):
            return dec(aff(msm(b"".join(scalars), gb * npts), 1))
        g2x = point([two], 1)
        assert point([one] * 4, 4) == point([four], 1) == point([two, two], 2)          # G+G+G+G == 2(2G)
        assert point([one] * 10, 10) == point([ten], 1)                                   # timesScalar(10) == 10 additions
        assert g2x is not None and point([one], 1) == G
        assert point([le(B.R_ORDER, 32)], 1) is None                                      # r G == infinity
        assert point([le(B.R_ORDER - 1, 32), one], 2) is None                             # (r - 1) G + G
        assert point([four, le(B.R_ORDER - 4, 32)], 2) is None                            # g?m_sub(4G, 4G) is zero
