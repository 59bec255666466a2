"""MSM pipeline index logic (msm_core.cuh bodies, run sequentially on the CPU by
tests/host/host_msm.cpp) against the oracle: signed digits with the carry window, counting
sort, task split, segment running sums, bit-plane sums, Horner -- for several window sizes
and task caps, on random and adversarial inputs (SURVEY.md 8d, C2 side set)."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import bn128 as O
from test_host_math import G, b, gen_points

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def host_msm():
    src = os.path.join(ROOT, "tests", "host", "host_msm.cpp")
    so = os.path.join(ROOT, "tests", "host", "libhost_msm.so")
    csrc = os.path.join(ROOT, "wasmsnark_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)


def adversarial(F, gen, rnd, n):
    """duplicates, P and -P, infinity entries, scalars {0, 1, r-1, r, 2^256-1, 2^255, ...}"""
    base = gen_points(F, gen, rnd, 8)
    pts, sc = [], []
    special = [0, 1, 2, O.R_ORDER - 1, O.R_ORDER, (1 << 256) - 1, 1 << 255, (1 << 255) - 1, (1 << 128) - 1, 0x8000, 0x7FFF, 0xFFFF, 0x10000]
    for i in range(n):
        k = rnd.randrange(10)
        P = base[rnd.randrange(len(base))]
        if k == 0:
            P = O.pt_zero(F)
        elif k == 1:
            P = O.pt_neg(F, P)
        pts.append(P)
        sc.append(special[rnd.randrange(len(special))] if rnd.randrange(3) else rnd.getrandbits(256))
    return sc, pts


@pytest.mark.parametrize("g", [1, 2])
def test_pipeline_matches_oracle(host_msm, host_math, g):
    rnd = random.Random(40 + g)
    H = G(host_math, g)
    F = H.F
    gen = O.G1_GEN if g == 1 else O.G2_GEN
    fn = host_msm.h_msm_g1 if g == 1 else host_msm.h_msm_g2
    cases = []
    n = 48 if g == 1 else 24
    pts = gen_points(F, gen, rnd, n)
    cases.append(([rnd.getrandbits(256) for _ in range(n)], pts))
    cases.append(([rnd.randrange(O.R_ORDER) for _ in range(n)], pts))
    cases.append(adversarial(F, gen, rnd, n))
    cases.append(([5] * n, [pts[0]] * n))                 # one bucket takes everything
    cases.append(([0] * 4, pts[:4]))                      # empty sum
    cases.append(([(1 << 256) - 1], pts[:1]))             # single pair, carry into the top window
    for sc, ps in cases:
        want = H.want(O.multiexp_naive(F, sc, ps))
        sb = b"".join(b(s) for s in sc)
        pb = b"".join(H.affine_bytes(P) for P in ps)
        for c, T in ((16, 64), (13, 4), (7, 3), (4, 64), (2, 2), (11, 1)):
            out = ctypes.create_string_buffer(3 * H.fs)
            fn(sb, pb, len(sc), c, T, out)
            got = tuple(H.fi(out.raw[k * H.fs:(k + 1) * H.fs]) for k in range(3))
            assert got == want, (c, T, len(sc))
        # window-table mode (resident keys): one shared bucket set, windows up to 20 bits
        fnt = host_msm.h_msm_g1_table if g == 1 else host_msm.h_msm_g2_table
        for c, T in ((20, 64), (16, 3), (9, 2), (17, 64)):
            out = ctypes.create_string_buffer(3 * H.fs)
            fnt(sb, pb, len(sc), c, T, out)
            got = tuple(H.fi(out.raw[k * H.fs:(k + 1) * H.fs]) for k in range(3))
            assert got == want, ("table", c, T, len(sc))
