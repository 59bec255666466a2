"""C5: MNT6753 wide-limb path on the GPU through the C ABI -- Fq3 products (1k pairs), G1 MSM on small sets
vs the definitional sum (incl. the reference's hard-coded P1 + P2 known answer) and n = 2^18 via the
closed form of sum s_i (P0 + i D).  The reference has no MSM for this curve (parity vs definitional sum)."""
import ctypes
import random

import pytest

from oracle import mnt6753 as M
from test_oracle_mnt import P1, P2, P12
from wasmsnark_b200._lib import as_ptr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bn():
    from wasmsnark_b200 import build_bn128
    b = build_bn128(0)
    yield b
    b.terminate()


def msm(bn, scalars, points):
    n = len(scalars)
    out = ctypes.create_string_buffer(288)
    sb = b"".join(s.to_bytes(96, "little") for s in scalars)
    pb = b"".join(M.affine_bytes_m(P) for P in points)
    bn.ctx.check(bn.lib.ws_mnt6753_g1_multiexp(bn.ctx.h, as_ptr(sb), as_ptr(pb), n, ctypes.cast(out, ctypes.c_void_p)))
    aff = ctypes.create_string_buffer(288)
    bn.ctx.check(bn.lib.ws_mnt6753_g1_affine(bn.ctx.h, ctypes.cast(out, ctypes.c_void_p), ctypes.cast(aff, ctypes.c_void_p), 1, 1))
    return tuple(int.from_bytes(aff.raw[96 * i:96 * i + 96], "little") for i in range(3))


def test_fq3_mul_1k(bn):
    rnd = random.Random(21)
    xs = [[rnd.randrange(M.Q) for _ in range(3)] for _ in range(1000)]
    ys = [[rnd.randrange(M.Q) for _ in range(3)] for _ in range(1000)]
    xs[0], ys[0] = [0, 0, 0], [1, 2, 3]
    xs[1], ys[1] = [M.Q - 1] * 3, [M.Q - 1] * 3
    enc = lambda vs: b"".join(M.to_mont(v).to_bytes(96, "little") for v in vs)
    X = b"".join(enc(v) for v in xs)
    Y = b"".join(enc(v) for v in ys)
    out = ctypes.create_string_buffer(288 * 1000)
    bn.ctx.check(bn.lib.ws_mnt6753_fq3_mul_batch(bn.ctx.h, as_ptr(X), as_ptr(Y), ctypes.cast(out, ctypes.c_void_p), 1000))
    raw = out.raw
    for i in range(1000):
        got = tuple(M.from_mont(int.from_bytes(raw[288 * i + 96 * k:288 * i + 96 * k + 96], "little")) for k in range(3))
        assert got == M.f3_mul(xs[i], ys[i]), i


def test_known_answer_and_small_sets(bn):
    assert msm(bn, [1, 1], [P1, P2]) == (P12[0], P12[1], 1)            # test/mnt6753.js:124-155
    assert msm(bn, [M.R_ORDER], [M.G1_GEN]) == (0, 1, 0)               # r*G = infinity, test/mnt6753.js:199-221
    rnd = random.Random(22)
    pts = [M.times_scalar(M.G1_GEN, rnd.randrange(1, M.R_ORDER)) for _ in range(12)]
    pts[3] = None
    pts[5] = M.neg(pts[4])
    sc = [rnd.getrandbits(768) if i % 3 else rnd.randrange(M.R_ORDER) for i in range(12)]
    sc[4] = sc[5] = 7
    sc[6], sc[7], sc[8] = 0, 1, (1 << 768) - 1
    assert msm(bn, sc, pts) == M.affine_result(M.multiexp_naive(sc, pts))
    assert msm(bn, [], []) == (0, 1, 0)


def test_msm_2p18_closed_form(bn):
    n = 1 << 18
    rnd = random.Random(23)
    P0 = M.times_scalar(M.G1_GEN, rnd.randrange(1, M.R_ORDER))
    D = M.times_scalar(M.G1_GEN, rnd.randrange(1, M.R_ORDER))
    Q = M.Q
    xs, ys = [P0[0]], [P0[1]]
    m = 1
    while len(xs) < n:                      # P_{i+m} = P_i + m*D, batched inversion
        mD = M.times_scalar(D, m)
        cnt = min(m, n - len(xs))
        dens = [(mD[0] - xs[i]) % Q for i in range(cnt)]
        pref = [1] * (cnt + 1)
        for i in range(cnt):
            pref[i + 1] = pref[i] * dens[i] % Q
        inv = pow(pref[cnt], -1, Q)
        xs.extend([None] * cnt)
        ys.extend([None] * cnt)
        for i in reversed(range(cnt)):
            di = inv * pref[i] % Q
            inv = inv * dens[i] % Q
            lam = (mD[1] - ys[i]) * di % Q
            x3 = (lam * lam - xs[i] - mD[0]) % Q
            xs[m + i] = x3
            ys[m + i] = (lam * (xs[i] - x3) - ys[i]) % Q
        m *= 2
    sc = [rnd.randrange(M.R_ORDER) for _ in range(n)]
    want = M.add(M.times_scalar(P0, sum(sc) % M.R_ORDER), M.times_scalar(D, sum(i * s for i, s in enumerate(sc)) % M.R_ORDER))
    out = ctypes.create_string_buffer(288)
    sb = b"".join(s.to_bytes(96, "little") for s in sc)
    pb = b"".join(M.to_mont(x).to_bytes(96, "little") + M.to_mont(y).to_bytes(96, "little") for x, y in zip(xs, ys))
    bn.ctx.check(bn.lib.ws_mnt6753_g1_multiexp(bn.ctx.h, as_ptr(sb), as_ptr(pb), n, ctypes.cast(out, ctypes.c_void_p)))
    aff = ctypes.create_string_buffer(288)
    bn.ctx.check(bn.lib.ws_mnt6753_g1_affine(bn.ctx.h, ctypes.cast(out, ctypes.c_void_p), ctypes.cast(aff, ctypes.c_void_p), 1, 1))
    got = tuple(int.from_bytes(aff.raw[96 * i:96 * i + 96], "little") for i in range(3))
    assert got == M.affine_result(want)
