"""Shared helpers for the parity tests: byte layouts, deterministic inputs, reference handles."""
import os
import random
import threading

from oracle import bn128 as O
from oracle import ref_wasm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q, RR = O.Q, O.R_ORDER
MNT_Q = 0x01C4C62D92C41110229022EEE2CDADB7F997505B8FAFED5EB7E8F96C97D87307FDB925E8A0ED8D99D124D9A15AF79DB26C5C28C859A99B3EEBCA9429212636B9DFF97634993AA4D6C381BC3F0057974EA099170FA13A4FD90776E240000001


def b32(x):
    return x.to_bytes(32, "little")


def i32(bs):
    return int.from_bytes(bs, "little")


def edge_values(p):
    """test/f1.js:299-344 edge grid of the reference, extended with limb boundaries."""
    h = (p - 1) // 2
    return [0, 1, 2, p - 1, p - 2, h, h + 1, h - 1, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, 1 << 64]


# ---- G1 point sets ---------------------------------------------------------------------
def g1_affine_m(P):
    """oracle Jacobian (plain) -> 64-byte affine Montgomery; infinity -> x = 0 (all zero)."""
    if O.pt_is_zero(O.Fq, P):
        return bytes(64)
    a = O.pt_affine(O.Fq, P)
    return b32(O.to_mont(a[0])) + b32(O.to_mont(a[1]))


def g2_affine_m(P):
    if O.pt_is_zero(O.Fq2, P):
        return bytes(128)
    a = O.pt_affine(O.Fq2, P)
    return b"".join(b32(O.to_mont(c)) for c in (a[0][0], a[0][1], a[1][0], a[1][1]))


def g1_result_tuple(b96):
    """96-byte Montgomery X|Y|Z -> plain ints."""
    return tuple(O.from_mont(i32(b96[k * 32:k * 32 + 32])) for k in range(3))


def g2_result_tuple(b192):
    v = [O.from_mont(i32(b192[k * 32:k * 32 + 32])) for k in range(6)]
    return ((v[0], v[1]), (v[2], v[3]), (v[4], v[5]))


def arith_points(F, gen, n, seed, offset=0):
    """P_i = P0 + i*D: one affine chord addition per point with a batched
    inversion, so large sets are cheap and sum(s_i P_i) has the closed form
    (sum s_i) P0 + (sum i s_i) D.  With `offset` the list starts at P0 + offset*D (a rank's slice of one global
    progression).  Returns (list of affine (x, y) plain, P0, D)."""
    rnd = random.Random(seed)
    base = (gen[0], gen[1], F.one)
    k0, k1 = rnd.randrange(1, RR), rnd.randrange(1, RR)
    P0 = O.pt_affine(F, O.pt_times_scalar(F, base, k0))
    D = O.pt_affine(F, O.pt_times_scalar(F, base, k1))
    assert F is O.Fq, "fast path written for G1"
    start = P0 if offset == 0 else O.pt_affine(F, O.pt_add(F, (P0[0], P0[1], 1), O.pt_times_scalar(F, (D[0], D[1], 1), offset)))
    xs, ys = [start[0]], [start[1]]
    dx, dy = D[0], D[1]
    # chunks of independent additions: P_{i+m} = P_i + m*D, batch-inverted
    m = 1
    while len(xs) < n:
        mD = O.pt_affine(F, O.pt_times_scalar(F, (dx, dy, 1), m))
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
            y3 = (lam * (xs[i] - x3) - ys[i]) % Q
            xs[m + i] = x3
            ys[m + i] = y3
        m *= 2
    return list(zip(xs[:n], ys[:n])), P0, D


def g1_points_bytes(aff):
    return b"".join(b32(O.to_mont(x)) + b32(O.to_mont(y)) for x, y in aff)


def closed_form_sum(scalars, P0, D):
    F = O.Fq
    s0 = sum(scalars) % RR
    s1 = sum(i * s for i, s in enumerate(scalars)) % RR
    a = O.pt_times_scalar(F, (P0[0], P0[1], 1), s0)
    bb = O.pt_times_scalar(F, (D[0], D[1], 1), s1)
    return O.pt_affine(F, O.pt_add(F, a, bb))


# ---- the reference's own code (compiled WASM image), optional ------------------------------
def have_ref():
    return ref_wasm.available()


class RefWorkerPool:
    """The reference's worker pool (src/bn128.js:173-265): W instances of its WASM image, created once
    (buildBn128), each on its own thread; g1_multiexp splits by contiguous ranges and the partial points are
    added with g1m_add on the main instance (src/bn128.js:353-383)."""

    def __init__(self, nthreads):
        from concurrent.futures import ThreadPoolExecutor
        self.W = max(1, nthreads)
        self.inst = [ref_wasm.RefInstance(pages=5000) for _ in range(self.W)]
        self.main = ref_wasm.RefInstance(pages=1000)
        self.p0, self.p1 = self.main.alloc(96), self.main.alloc(96)
        self.ex = ThreadPoolExecutor(max_workers=self.W)

    def g1_multiexp_affine(self, scalars, points):
        n = len(scalars) // 32
        W = max(1, min(self.W, n))
        per = n // W

        def work(i):
            cnt = per if i < W - 1 else n - per * (W - 1)
            return self.inst[i].cmd_g1_multiexp(scalars[i * per * 32:(i * per + cnt) * 32], points[i * per * 64:(i * per + cnt) * 64], cnt)
        res = list(self.ex.map(work, range(W)))
        m = self.main
        m.call("g1m_zero", self.p0)
        for r in res:
            m.put(self.p1, r)
            m.call("g1m_add", self.p0, self.p1, self.p0)
        m.call("g1m_affine", self.p0, self.p0)
        return m.get(self.p0, 96)

    def close(self):
        self.ex.shutdown()
        for i in self.inst:
            i.close()
        self.main.close()


def ref_g1_multiexp_threads(scalars, points, nthreads=None):
    """One-shot form of RefWorkerPool.g1_multiexp_affine (pool built and torn down inside)."""
    n = len(scalars) // 32
    pool = RefWorkerPool(max(1, min(nthreads or min(os.cpu_count() or 1, 16), n)))
    try:
        return pool.g1_multiexp_affine(scalars, points)
    finally:
        pool.close()
