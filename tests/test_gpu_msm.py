"""C2 multi-scalar multiplication parity on the GPU, through the C ABI.

 * small sets vs the definitional oracle sum (G1 and G2), every window size that changes the
   carry-window shape, forced task splits (warp combine path);
 * adversarial set: duplicates, P / -P, infinity (x = 0) entries, scalars {0,1,r-1,r,2^256-1,..};
 * 2^14 prefix vs the reference's own g1m_multiexp2 (w = 7), after g1m_affine as
   test/bn128_prover.js:31-46 compares;
 * full size n = 2^20 (BASELINE.json configs[1]): raw 256-bit scalars and scalars < r, checked
   with the closed form of sum s_i (P0 + i D) and, when oracle/_ref is present, against the
   reference algorithm run over the host cores.
"""
import ctypes
import random

import pytest

from oracle import bn128 as O
from oracle import ref_wasm
import util
from util import RR, b32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bn():
    from wasmsnark_b200 import build_bn128
    b = build_bn128(0)
    yield b
    b.terminate()


def dev_multiexp(bn, g, scalars_b, points_b, n, c=0, T=0):
    """call the *_dev entry with explicit window bits (torch only allocates device memory); stream NULL = the ctx stream."""
    import torch
    psize, osize = (64, 96) if g == 1 else (128, 192)
    ds = torch.frombuffer(bytearray(scalars_b), dtype=torch.uint8).cuda()
    dp = torch.frombuffer(bytearray(points_b), dtype=torch.uint8).cuda()
    dout = torch.zeros(256, dtype=torch.uint8, device="cuda")
    fn = bn.lib.ws_bn128_g1_multiexp_dev if g == 1 else bn.lib.ws_bn128_g2_multiexp_dev
    bn.lib.ws_set_msm_task_cap(bn.ctx.h, T)
    torch.cuda.synchronize()
    bn.ctx.check(fn(bn.ctx.h, ds.data_ptr(), dp.data_ptr(), n, dout.data_ptr(), c, None))
    torch.cuda.synchronize()
    bn.lib.ws_set_msm_task_cap(bn.ctx.h, 0)
    return bytes(dout[:osize].cpu().numpy())


def small_cases(F, gen, rnd, n):
    from test_host_math import gen_points
    from test_host_msm import adversarial
    pts = gen_points(F, gen, rnd, n)
    cases = [
        ([rnd.getrandbits(256) for _ in range(n)], pts),
        ([rnd.randrange(RR) for _ in range(n)], pts),
        adversarial(F, gen, rnd, n),
        ([5] * n, [pts[0]] * n),
        ([0] * 4, pts[:4]),
        ([(1 << 256) - 1], pts[:1]),
        ([], []),
    ]
    return cases


@pytest.mark.parametrize("g", [1, 2])
def test_small_sets_vs_oracle(bn, g):
    rnd = random.Random(100 + g)
    F = O.Fq if g == 1 else O.Fq2
    gen = O.G1_GEN if g == 1 else O.G2_GEN
    enc = util.g1_affine_m if g == 1 else util.g2_affine_m
    dec = util.g1_result_tuple if g == 1 else util.g2_result_tuple
    for sc, ps in small_cases(F, gen, rnd, 40 if g == 1 else 20):
        want = O.pt_affine(F, O.multiexp_naive(F, sc, ps))
        sb = b"".join(b32(s) for s in sc)
        pb = b"".join(enc(P) for P in ps)
        aff = bn.g1_affine if g == 1 else bn.g2_affine
        host = bn.g1_multiexp(sb, pb) if g == 1 else bn.g2_multiexp(sb, pb)
        assert dec(aff(host)) == want
        if not sc:
            continue
        for c, T in ((16, 0), (13, 4), (7, 3), (4, 0), (2, 2), (11, 1)):
            assert dec(aff(dev_multiexp(bn, g, sb, pb, len(sc), c, T))) == want, (c, T, len(sc))


@pytest.mark.parametrize("g", [1, 2])
def test_adversarial_4096_and_independent_points(bn, g):
    """SURVEY 8d C2 side sets: n = 4096 drawn from 8 base points with duplicates, P / -P pairs, infinity entries and
    the scalars {0, 1, r-1, r, 2^256-1, ...} (every bucket sees repeated and cancelling points); and a set of
    independent k_i * G points (no arithmetic-progression structure).  Oracle: its own Pippenger."""
    from test_host_math import gen_points
    from test_host_msm import adversarial
    rnd = random.Random(4096 + g)
    F = O.Fq if g == 1 else O.Fq2
    gen = O.G1_GEN if g == 1 else O.G2_GEN
    enc = util.g1_affine_m if g == 1 else util.g2_affine_m
    dec = util.g1_result_tuple if g == 1 else util.g2_result_tuple
    aff = bn.g1_affine if g == 1 else bn.g2_affine
    fn = bn.g1_multiexp if g == 1 else bn.g2_multiexp
    n_ind = 1536 if g == 1 else 384
    ind = gen_points(F, gen, rnd, n_ind)
    for sc, ps in (adversarial(F, gen, rnd, 4096 if g == 1 else 1024), ([rnd.getrandbits(256) for _ in range(n_ind)], ind)):
        want = O.pt_affine(F, O.multiexp(F, sc, ps))
        sb = b"".join(b32(s) for s in sc)
        pb = b"".join(enc(P) for P in ps)
        assert dec(aff(fn(sb, pb))) == want
        assert dec(aff(dev_multiexp(bn, g, sb, pb, len(sc), 9, 5))) == want


def test_group_identities_from_reference_suite(bn):
    """test/bn128.js:84-185 through the MSM entry points: r * G = infinity, 10 * P = P added ten times,
    P + (-P) = infinity, for G1 and G2."""
    for g, F, gen, enc, dec, fn, aff in ((1, O.Fq, O.G1_GEN, util.g1_affine_m, util.g1_result_tuple, bn.g1_multiexp, bn.g1_affine),
                                         (2, O.Fq2, O.G2_GEN, util.g2_affine_m, util.g2_result_tuple, bn.g2_multiexp, bn.g2_affine)):
        G_ = (gen[0], gen[1], F.one)
        inf = O.pt_affine(F, O.pt_zero(F))
        assert dec(aff(fn(b32(RR), enc(G_)))) == inf
        assert dec(aff(fn(b32(10), enc(G_)))) == dec(aff(fn(b32(1) * 10, enc(G_) * 10)))
        assert dec(aff(fn(b32(7) + b32(7), enc(G_) + enc(O.pt_neg(F, G_))))) == inf
        assert dec(aff(fn(b32(RR - 1), enc(G_)))) == O.pt_affine(F, O.pt_neg(F, G_))


def test_back_to_back_async_multiexps_without_a_sync_between_them(bn):
    """Six ws_bn128_g?_multiexp_dev calls (G1 and G2 alternating, different sizes, scalars and output buffers) queued on ONE
    context with no synchronisation between them: they share the context's MSM workspace (digits, bucket tables, partial
    sums), so each must be ordered behind its predecessor on the device -- every result against the oracle's sum.  Under the
    emulator's deferred modes (CUEMU_ASYNC) a missing dependency between consecutive calls shows here."""
    import torch
    from test_host_math import gen_points
    rnd = random.Random(23)
    jobs = []
    for i, n in enumerate((300, 48, 700, 17, 512, 96)):
        g = 1 + (i & 1)
        F, gen, enc = (O.Fq, O.G1_GEN, util.g1_affine_m) if g == 1 else (O.Fq2, O.G2_GEN, util.g2_affine_m)
        ps = gen_points(F, gen, rnd, n)
        sc = [rnd.getrandbits(256) for _ in range(n)]
        jobs.append((g, n, torch.frombuffer(bytearray(b"".join(b32(v) for v in sc)), dtype=torch.uint8).cuda(),
                     torch.frombuffer(bytearray(b"".join(enc(P) for P in ps)), dtype=torch.uint8).cuda(), torch.zeros(256, dtype=torch.uint8, device="cuda"),
                     O.pt_affine(F, O.multiexp(F, sc, ps))))
    torch.cuda.synchronize()
    for g, n, ds, dp, dout, _ in jobs:
        fn = bn.lib.ws_bn128_g1_multiexp_dev if g == 1 else bn.lib.ws_bn128_g2_multiexp_dev
        bn.ctx.check(fn(bn.ctx.h, ds.data_ptr(), dp.data_ptr(), n, dout.data_ptr(), 0, None))
    torch.cuda.synchronize()
    for g, n, ds, dp, dout, want in jobs:
        out = bytes(dout[:96 if g == 1 else 192].cpu().numpy())
        got = (util.g1_result_tuple(bn.g1_affine(out)) if g == 1 else util.g2_result_tuple(bn.g2_affine(out)))
        assert got == want, (g, n)


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not built")
def test_from_montgomery_methods_equal_the_reference_functions(bn):
    """Bn128.g1_fromMontgomery / g2_fromMontgomery (src/bn128.js:429-439): un-normalised multiexp results (Z != 1) and the
    point at infinity through the mirror's methods and through the reference image's g1m_ / g2m_fromMontgomery, byte for byte;
    g?_affine followed by g?_fromMontgomery equals the combined g?_affine_std the prover's output path uses."""
    rnd = random.Random(31)
    inst = ref_wasm.RefInstance(pages=100)
    for g, gen, enc, fn, size, ref_fn, mexp, aff, aff_std in ((1, O.G1_GEN, util.g1_affine_m, bn.g1_fromMontgomery, 96, "g1m_fromMontgomery", bn.g1_multiexp, bn.g1_affine, bn.g1_affine_std),
                                                               (2, O.G2_GEN, util.g2_affine_m, bn.g2_fromMontgomery, 192, "g2m_fromMontgomery", bn.g2_multiexp, bn.g2_affine, bn.g2_affine_std)):
        F = O.Fq if g == 1 else O.Fq2
        G_ = (gen[0], gen[1], F.one)
        pts = [mexp(b32(rnd.getrandbits(256)) + b32(rnd.getrandbits(200)), enc(G_) * 2) for _ in range(4)] + [mexp(b32(RR), enc(G_))]
        px = inst.alloc(size)
        for P in pts:
            inst.put(px, P)
            inst.call(ref_fn, px, px)
            assert fn(P) == inst.get(px, size)
            assert fn(aff(P)) == aff_std(P)
        assert fn(b"".join(pts)) == b"".join(fn(P) for P in pts)
    inst.close()


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not built")
def test_2p14_vs_reference_multiexp2(bn):
    n = 1 << 14
    aff, P0, D = util.arith_points(O.Fq, O.G1_GEN, n, 0xB202)
    rnd = random.Random(0xB201)
    sc = [rnd.getrandbits(256) for _ in range(n)]
    for k in range(0, n, 97):          # sprinkle zeros / ones / small values like a real witness
        sc[k] = (0, 1, rnd.getrandbits(32))[k % 3]
    sb = b"".join(b32(s) for s in sc)
    pb = bytearray(util.g1_points_bytes(aff))
    for k in range(5, n, 251):         # some infinity entries (x = 0)
        pb[64 * k:64 * k + 64] = bytes(64)
    pb = bytes(pb)
    want = util.ref_g1_multiexp_threads(sb, pb, nthreads=8)
    assert bn.g1_affine(bn.g1_multiexp(sb, pb)) == want


@pytest.mark.parametrize("reduced", [False, True])
def test_full_size_2p20(bn, reduced):
    n = 1 << 20
    aff, P0, D = util.arith_points(O.Fq, O.G1_GEN, n, 0xB202)
    rnd = random.Random(0xB201 + reduced)
    sc = [rnd.randrange(RR) if reduced else rnd.getrandbits(256) for _ in range(n)]
    sb = b"".join(b32(s) for s in sc)
    pb = util.g1_points_bytes(aff)
    got = bn.g1_affine(bn.g1_multiexp(sb, pb))
    assert util.g1_result_tuple(got) == util.closed_form_sum(sc, P0, D)
    if ref_wasm.available() and not reduced:
        assert got == util.ref_g1_multiexp_threads(sb, pb)


@pytest.mark.parametrize("g,table_c", [(1, 16), (1, 20), (1, 9), (2, 12), (2, 17)])
def test_window_table_msm(bn, g, table_c):
    """Window-table mode (what resident proving keys use): same sums as the table-free path / the oracle."""
    import torch
    rnd = random.Random(300 + g + table_c)
    if g == 1:
        n = 5000
        aff, P0, D = util.arith_points(O.Fq, O.G1_GEN, n, 77)
        pb = bytearray(util.g1_points_bytes(aff))
        for k in range(3, n, 97):
            pb[64 * k:64 * k + 64] = bytes(64)               # infinity entries
        psize, osize = 64, 96
    else:
        from test_host_math import gen_points
        n = 60
        pts = gen_points(O.Fq2, O.G2_GEN, rnd, n)
        pts[7] = O.pt_zero(O.Fq2)
        pb = bytearray(b"".join(util.g2_affine_m(P) for P in pts))
        psize, osize = 128, 192
    sc = [rnd.getrandbits(256) if i % 5 else (0, 1, (1 << 256) - 1, O.R_ORDER, rnd.getrandbits(32))[(i // 5) % 5] for i in range(n)]
    sb = b"".join(b32(s) for s in sc)
    want = bn.g1_affine(bn.g1_multiexp(sb, bytes(pb))) if g == 1 else bn.g2_affine(bn.g2_multiexp(sb, bytes(pb)))
    W = bn.lib.ws_msm_table_windows(table_c)
    assert W == (257 + table_c - 1) // table_c
    ds = torch.frombuffer(bytearray(sb), dtype=torch.uint8).cuda()
    dp = torch.frombuffer(pb, dtype=torch.uint8).cuda()
    dt = torch.zeros(W * n * psize, dtype=torch.uint8, device="cuda")
    dout = torch.zeros(256, dtype=torch.uint8, device="cuda")
    build = bn.lib.ws_bn128_g1_build_table_dev if g == 1 else bn.lib.ws_bn128_g2_build_table_dev
    run = bn.lib.ws_bn128_g1_multiexp_table_dev if g == 1 else bn.lib.ws_bn128_g2_multiexp_table_dev
    torch.cuda.synchronize()
    bn.ctx.check(build(bn.ctx.h, dp.data_ptr(), n, table_c, dt.data_ptr(), None))
    bn.ctx.check(run(bn.ctx.h, ds.data_ptr(), dt.data_ptr(), n, table_c, dout.data_ptr(), None))
    bn.ctx.timings()
    torch.cuda.synchronize()
    got = bytes(dout[:osize].cpu().numpy())
    got = bn.g1_affine(got) if g == 1 else bn.g2_affine(got)
    assert got == want
    if g == 1:          # the table itself: row w holds 2^(c w) * P_i
        tab = bytes(dt[:2 * n * 64].cpu().numpy())
        i = 11
        x = O.from_mont(int.from_bytes(tab[64 * (n + i):64 * (n + i) + 32], "little"))
        y = O.from_mont(int.from_bytes(tab[64 * (n + i) + 32:64 * (n + i) + 64], "little"))
        want_pt = O.pt_affine(O.Fq, O.pt_times_scalar(O.Fq, (aff[i][0], aff[i][1], 1), 1 << table_c))
        assert (x, y) == (want_pt[0], want_pt[1])
