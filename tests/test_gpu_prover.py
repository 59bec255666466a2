"""C3 (NTT) and C4 (H pipeline + full Groth16 prove) parity on the GPU, through the C ABI.

Golden data comes from the reference's own code (tests/golden/make_golden.py), including the shipped
example/bn128/proof_good0.json (r = s = 0).  Full sizes are checked through size-independent
properties: fft -> ifft round trip at 2^22 (BASELINE configs[2]) and the even/odd interleave identity
of test/fft.js:71-121."""
import json
import os
import random

import numpy as np
import pytest

from oracle import bn128 as O
from oracle import ref_wasm

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
EXAMPLE = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "example_bn128")


@pytest.fixture(scope="module")
def bn():
    from wasmsnark_b200 import build_bn128
    b = build_bn128(0)
    yield b
    b.terminate()


def mont_bytes(vals):
    return b"".join(O.to_mont(v, O.R_ORDER).to_bytes(32, "little") for v in vals)


def from_mont_list(bs):
    return [O.from_mont(int.from_bytes(bs[i:i + 32], "little"), O.R_ORDER) for i in range(0, len(bs), 32)]


def test_fft_golden_vectors(bn):
    for v in json.load(open(os.path.join(GOLD, "fft_vectors.json"))):
        x = bytes.fromhex(v["in"])
        got = bn.fft(x, v["odd"]) if v["fn"] == "fft_fft" else bn.ifft(x, v["odd"])
        assert got.hex() == v["out"], (v["fn"], v["n"], v["odd"])


@pytest.mark.parametrize("logn", [1, 2, 3, 4, 5, 6, 7, 9, 10, 12])
def test_fft_vs_oracle_all_radix_mixes(bn, logn):
    rnd = random.Random(logn)
    n = 1 << logn
    x = [rnd.randrange(O.R_ORDER) for _ in range(n)]
    xb = mont_bytes(x)
    for odd in (0, 1):
        assert from_mont_list(bn.fft(xb, odd)) == O.fft(x, odd)
        assert from_mont_list(bn.ifft(xb, odd)) == O.ifft(x, odd)


def test_fft_reference_zeros_pattern(bn):
    """test/fft.js:53-70 ("test with zeros"), step for step: 0 / 1 alternating, N = 1024, fft_toMontgomeryN -> fft_ifft ->
    fft_fft -> fft_fromMontgomeryN gives the input back; the transform of the pattern itself also equals the oracle's."""
    from wasmsnark_b200 import _lib
    n = 1024
    vals = [i & 1 for i in range(n)]
    x = b"".join(v.to_bytes(32, "little") for v in vals)
    xm = bn.fp_batch(_lib.FIELD_BN128_FR, _lib.OP_TOMONT, x)
    mid = bn.ifft(xm, 0)
    assert from_mont_list(mid) == O.ifft(vals, 0)
    back = bn.fp_batch(_lib.FIELD_BN128_FR, _lib.OP_FROMMONT, bn.fft(mid, 0))
    assert back == x


def test_fft_rejects_bad_sizes(bn):
    from wasmsnark_b200._lib import WsError
    for n in (3, 12, 1000):
        with pytest.raises(WsError):
            bn.fft(bytes(32 * n))


@pytest.mark.parametrize("logn", [16, 20])
def test_fft_interleave_identity(bn, logn):
    """test/fft.js:71-121 at N = 2^16 and 2^20: FFT_2N(x | 0) = interleave(FFT_N(x), FFT_N(x, odd))."""
    n = 1 << logn
    rng = np.random.Generator(np.random.PCG64(0xB203))
    raw = np.frombuffer(rng.bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    raw[:, 31] &= 0x1F                                       # < 2^253 < r: canonical residues
    x = raw.tobytes()
    big = np.frombuffer(bn.fft(x + bytes(32 * n)), dtype=np.uint8).reshape(2 * n, 32)
    even = np.frombuffer(bn.fft(x, 0), dtype=np.uint8).reshape(n, 32)
    odd = np.frombuffer(bn.fft(x, 1), dtype=np.uint8).reshape(n, 32)
    assert (big[0::2] == even).all() and (big[1::2] == odd).all()


def test_fft_round_trip_2p22(bn):
    """BASELINE configs[2]: size 2^22 forward + inverse returns the input bit-exactly."""
    n = 1 << 22
    rng = np.random.Generator(np.random.PCG64(0xB203))
    raw = np.frombuffer(rng.bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    raw[:, 31] &= 0x1F                      # < 2^253 < r: canonical elements
    x = raw.tobytes()
    y = bn.fft(x)
    assert y != x
    assert bn.ifft(y) == x
    # spot check against the oracle on a 2^12 prefix problem
    m = 1 << 12
    xs = [int.from_bytes(x[32 * i:32 * i + 32], "little") for i in range(m)]
    got = bn.fft(x[:32 * m], 1)
    want = [v * (1 << 256) % O.R_ORDER for v in O.fft([O.from_mont(v, O.R_ORDER) for v in xs], 1)]
    assert [int.from_bytes(got[32 * i:32 * i + 32], "little") for i in range(m)] == want


def test_synth_small_calch_and_proofs(bn):
    pk = open(os.path.join(GOLD, "synth_small.pkey.bin"), "rb").read()
    wit = open(os.path.join(GOLD, "synth_small.witness.bin"), "rb").read()
    g = json.load(open(os.path.join(GOLD, "synth_small.golden.json")))
    k = O.parse_pkey(pk)
    h = bn.calcH(wit, pk[k["pPolsA"]:k["pPolsB"]], pk[k["pPolsB"]:k["pPointsA"]], k["nVars"], k["domainSize"])
    assert h.hex() == g["calcH"]
    assert bn.groth16GenProof(wit, pk, r=bytes(32), s=bytes(32)) == g["proof_r0s0"]
    r, s = bytes.fromhex(g["r"]), bytes.fromhex(g["s"])
    assert bn.groth16GenProof(wit, pk, r=r, s=s) == g["proof_rs"]
    key = bn.load_proving_key(pk)                        # resident key with window tables (default)
    assert (key.nVars, key.nPublic, key.domainSize) == (300, 5, 256)
    for _ in range(2):                                   # repeated proofs
        assert bn.groth16GenProof(wit, key, r=r, s=s) == g["proof_rs"]
    key.free()
    bn.lib.ws_set_key_tables(bn.ctx.h, 0)                # and without tables
    key = bn.load_proving_key(pk)
    assert bn.groth16GenProof(wit, key, r=r, s=s) == g["proof_rs"]
    key.free()
    bn.lib.ws_set_key_tables(bn.ctx.h, 1)


def test_h_priority_hook_gives_the_same_proof(bn):
    """ws_set_h_priority only moves the H chain to another stream: same bytes."""
    pk = open(os.path.join(GOLD, "synth_small.pkey.bin"), "rb").read()
    wit = open(os.path.join(GOLD, "synth_small.witness.bin"), "rb").read()
    g = json.load(open(os.path.join(GOLD, "synth_small.golden.json")))
    r, s = bytes.fromhex(g["r"]), bytes.fromhex(g["s"])
    key = bn.load_proving_key(pk)
    try:
        bn.lib.ws_set_h_priority(bn.ctx.h, 1)
        for _ in range(3):
            assert bn.groth16GenProof(wit, key, r=r, s=s) == g["proof_rs"]
    finally:
        bn.lib.ws_set_h_priority(bn.ctx.h, 0)
        key.free()


def test_stateless_call_key_cache(bn):
    """ws_bn128_groth16_prove keeps the previous call's uploaded key, keyed by content: same key -> reused, a key of
    the same length with different content -> replaced (never a stale proof), cache off -> same bytes."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import gen_synth_key
    pk = open(os.path.join(GOLD, "synth_small.pkey.bin"), "rb").read()
    wit = open(os.path.join(GOLD, "synth_small.witness.bin"), "rb").read()
    g = json.load(open(os.path.join(GOLD, "synth_small.golden.json")))
    r, s = bytes.fromhex(g["r"]), bytes.fromhex(g["s"])
    k = O.parse_pkey(pk)
    pk2, wit2 = gen_synth_key.make(k["nVars"], k["nPublic"], k["domainSize"], seed=77)      # another key
    want2 = O.groth16_gen_proof(wit2, pk2, int.from_bytes(r, "little"), int.from_bytes(s, "little"))
    a = k["pPointsA"]                                                                         # same length, two A points swapped
    pk3 = pk[:a + 64] + pk[a + 128:a + 192] + pk[a + 64:a + 128] + pk[a + 192:]
    assert len(pk3) == len(pk) and pk3 != pk
    want3 = O.groth16_gen_proof(wit, pk3, int.from_bytes(r, "little"), int.from_bytes(s, "little"))
    assert want3 != g["proof_rs"]
    for _ in range(2):
        assert bn.groth16GenProof(wit, pk, r=r, s=s) == g["proof_rs"]
        assert bn.groth16GenProof(wit, pk, r=r, s=s) == g["proof_rs"]
        p2 = bn.groth16GenProof(wit2, pk2, r=r, s=s)
        assert p2 != g["proof_rs"]
        assert p2 == want2
        assert bn.groth16GenProof(wit, pk3, r=r, s=s) == want3
    bn.lib.ws_set_key_cache(bn.ctx.h, 0)
    assert bn.groth16GenProof(wit2, pk2, r=r, s=s) == p2
    assert bn.groth16GenProof(wit, pk, r=r, s=s) == g["proof_rs"]
    bn.lib.ws_set_key_cache(bn.ctx.h, 1)


def test_random_rs_differs_and_bad_keys_rejected(bn):
    from wasmsnark_b200._lib import WsError
    pk = open(os.path.join(GOLD, "synth_small.pkey.bin"), "rb").read()
    wit = open(os.path.join(GOLD, "synth_small.witness.bin"), "rb").read()
    p1, p2 = bn.groth16GenProof(wit, pk), bn.groth16GenProof(wit, pk)      # CSPRNG r, s (src/bn128.js:642-664)
    assert p1 != p2
    with pytest.raises(WsError):
        bn.groth16GenProof(wit, pk[:1000])
    with pytest.raises(WsError):
        bn.groth16GenProof(wit[:64], pk)
    bad = bytearray(pk)
    bad[8:12] = (255).to_bytes(4, "little")              # domainSize not a power of two
    with pytest.raises(WsError):
        bn.groth16GenProof(wit, bytes(bad))


@pytest.mark.skipif(not os.path.exists(os.path.join(EXAMPLE, "proving_key.bin")), reason="example key not staged (oracle/build_ref.py)")
def test_example_key_golden_proof(bn):
    """config C4: example/bn128 proving_key.bin + witness.bin, r = s = 0 => proof_good0.json, bit for bit;
    and a fixed (r, s) against the proof the reference's own code produced for it."""
    pk = open(os.path.join(EXAMPLE, "proving_key.bin"), "rb").read()
    wit = open(os.path.join(EXAMPLE, "witness.bin"), "rb").read()
    good = json.load(open(os.path.join(EXAMPLE, "proof_good0.json")))
    gold = json.load(open(os.path.join(GOLD, "example_bn128.golden.json")))
    key = bn.load_proving_key(pk)
    proof = bn.groth16GenProof(wit, key, r=bytes(32), s=bytes(32))
    for kk in ("pi_a", "pi_b", "pi_c"):
        assert proof[kk] == good[kk] == gold[kk]
    proof = bn.groth16GenProof(wit, key, r=bytes.fromhex(gold["r"]), s=bytes.fromhex(gold["s"]))
    assert proof == gold["proof_rs"]
    key.free()
    assert bn.groth16GenProof(wit, pk, r=bytes(32), s=bytes(32)) == {kk: good[kk] for kk in ("pi_a", "pi_b", "pi_c")}


def test_sharded_partials_emulated_on_one_gpu(bn):
    """The N > 1 data path without N GPUs: the key loaded as shard 0/3, 1/3, 2/3 on one context, the three
    768-byte partial records concatenated as an all_gather would, then finish -- must equal the golden proof."""
    import ctypes
    import torch
    from wasmsnark_b200._lib import as_ptr
    pk = open(os.path.join(GOLD, "synth_small.pkey.bin"), "rb").read()
    wit = open(os.path.join(GOLD, "synth_small.witness.bin"), "rb").read()
    g = json.load(open(os.path.join(GOLD, "synth_small.golden.json")))
    world = 3
    d_w = torch.frombuffer(bytearray(wit), dtype=torch.uint8).cuda()
    d_rs = torch.frombuffer(bytearray(bytes.fromhex(g["r"]) + bytes.fromhex(g["s"])), dtype=torch.uint8).cuda()
    d_all = torch.zeros(768 * world, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(384, dtype=torch.uint8, device="cuda")
    keys = []
    for r in range(world):
        k = ctypes.c_void_p()
        bn.ctx.check(bn.lib.ws_bn128_pkey_load_shard(bn.ctx.h, as_ptr(pk), len(pk), r, world, ctypes.byref(k)))
        keys.append(k)
        torch.cuda.synchronize()
        bn.ctx.check(bn.lib.ws_bn128_groth16_partials_dev(bn.ctx.h, k, d_w.data_ptr(), d_all[768 * r:].data_ptr(), None))
    bn.ctx.check(bn.lib.ws_bn128_groth16_finish_dev(bn.ctx.h, keys[0], d_all.data_ptr(), world, d_rs.data_ptr(), d_out.data_ptr(), None))
    bn.ctx.timings()            # synchronises the ctx stream
    torch.cuda.synchronize()
    out = bytes(d_out.cpu().numpy())
    proof = {"pi_a": bn.bin2g1(out[:96]), "pi_b": bn.bin2g2(out[96:288]), "pi_c": bn.bin2g1(out[288:])}
    assert proof == g["proof_rs"]
    for k in keys:
        bn.lib.ws_bn128_pkey_free(k)


@pytest.mark.skipif(not (ref_wasm.available() and os.path.exists(os.path.join(EXAMPLE, "verification_key.json"))), reason="oracle/_ref not staged")
def test_random_rs_proof_accepted_by_reference_verifier(bn):
    """The reference's own end-to-end acceptance test (test/bn128_prover.js:51-80): a proof with CSPRNG r, s on
    test/data/{witness,proving_key}.bin (identical to example/bn128/) is accepted iff
    groth16Verify(verification_key, public, proof) === true -- here with the reference's verifier run from its
    compiled WASM image."""
    pk = open(os.path.join(EXAMPLE, "proving_key.bin"), "rb").read()
    wit = open(os.path.join(EXAMPLE, "witness.bin"), "rb").read()
    vk = json.load(open(os.path.join(EXAMPLE, "verification_key.json")))
    pub = json.load(open(os.path.join(EXAMPLE, "public.json")))
    key = bn.load_proving_key(pk)
    p1 = bn.groth16GenProof(wit, key)
    p2 = bn.groth16GenProof(wit, key)
    assert p1 != p2
    assert ref_wasm.groth16_verify(vk, pub, p1) and ref_wasm.groth16_verify(vk, pub, p2)
    assert bn.groth16Verify(vk, pub, p1) and bn.groth16Verify(vk, pub, p2)        # and by this library's verifier
    p1["pi_a"][1] = str((int(p1["pi_a"][1]) + 1) % O.Q)
    assert not ref_wasm.groth16_verify(vk, pub, p1)
    assert not bn.groth16Verify(vk, pub, p1)
    key.free()


# ---------------------------------------------------------------------------------------------------- round 2
def _synth_small():
    pk = open(os.path.join(GOLD, "synth_small.pkey.bin"), "rb").read()
    wit = open(os.path.join(GOLD, "synth_small.witness.bin"), "rb").read()
    g = json.load(open(os.path.join(GOLD, "synth_small.golden.json")))
    return pk, wit, g


@pytest.mark.parametrize("logn", [3, 6, 7, 8, 9, 13, 15, 18])
def test_fft_in_place_pass_counts_and_variants(bn, logn):
    """Every pass schedule of ntt_launch lands in the caller's buffer without a trailing copy (odd and even pass counts,
    one and two scratch buffers), and the occupancy variants of the radix-64 pass compute the same bytes."""
    n = 1 << logn
    rng = np.random.Generator(np.random.PCG64(logn))
    raw = np.frombuffer(rng.bytes(32 * n), dtype=np.uint8).reshape(n, 32).copy()
    raw[:, 31] &= 0x1F
    x = raw.tobytes()
    base = {}
    try:
        for variant in (0, 1, 2, 3, 4, 5, 6, 7):
            bn.lib.ws_set_ntt_variant(bn.ctx.h, variant)
            for odd in (0, 1):
                y = bn.fft(x, odd)
                assert bn.ifft(bn.fft(x, 0), 0) == x
                if variant == 0:
                    base[odd] = y
                    if logn <= 9:
                        xs = [O.from_mont(int.from_bytes(x[32 * i:32 * i + 32], "little"), O.R_ORDER) for i in range(n)]
                        assert from_mont_list(y) == O.fft(xs, odd)
                else:
                    assert y == base[odd]
    finally:
        bn.lib.ws_set_ntt_variant(bn.ctx.h, 0)


def test_fft_2p28_odd_is_rejected_not_garbage(bn):
    """BN128 Fr has 2-adicity 28: fft(n = 2^28, odd) would need a 2^29-th root of unity.  The size check
    comes before any allocation."""
    import ctypes
    from wasmsnark_b200._lib import WsError
    with pytest.raises(WsError):
        bn.ctx.check(bn.lib.ws_bn128_fr_fft_dev(bn.ctx.h, ctypes.c_void_p(256), 1 << 28, 1, 0, None))


def test_split_prover_entry_points_equal_the_single_call(bn):
    """witness_msms_dev + h_dev + h_msm_dev (the strong-scaling split: H computed once and handed over) on shards 0/2 and
    1/2 of one GPU, h copied slice-wise like the NCCL / peer-copy hand-over would, against the golden proof."""
    import ctypes
    import torch
    from wasmsnark_b200._lib import as_ptr
    pk, wit, g = _synth_small()
    world = 2
    d_w = torch.frombuffer(bytearray(wit), dtype=torch.uint8).cuda()
    d_rs = torch.frombuffer(bytearray(bytes.fromhex(g["r"]) + bytes.fromhex(g["s"])), dtype=torch.uint8).cuda()
    d_all = torch.zeros(768 * world, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(384, dtype=torch.uint8, device="cuda")
    keys, ranges = [], []
    for r in range(world):
        k = ctypes.c_void_p()
        bn.ctx.check(bn.lib.ws_bn128_pkey_load_shard(bn.ctx.h, as_ptr(pk), len(pk), r, world, ctypes.byref(k)))
        rg = (ctypes.c_uint32 * 6)()
        bn.lib.ws_bn128_pkey_shard_info(k, rg)
        keys.append(k)
        ranges.append(tuple(rg))
    domain = 256
    assert [rg[4:6] for rg in ranges] == [(0, 128), (128, 128)]
    d_h = torch.zeros(domain * 32, dtype=torch.uint8, device="cuda")
    bn.ctx.check(bn.lib.ws_bn128_groth16_h_dev(bn.ctx.h, keys[1], d_w.data_ptr(), d_h.data_ptr(), None))      # rank 1 owns H
    torch.cuda.synchronize()
    bn.ctx.timings()
    k = O.parse_pkey(pk)
    want_h = bn.calcH(wit, pk[k["pPolsA"]:k["pPolsB"]], pk[k["pPolsB"]:k["pPointsA"]], k["nVars"], k["domainSize"])
    assert bytes(d_h.cpu().numpy()) == want_h
    for r in range(world):
        # a non-owner only holds the hull of its witness ranges and its slice of h
        from wasmsnark_b200.parallel import witness_hull
        lo, cnt = witness_hull(ranges[r])
        w_r = torch.zeros_like(d_w)
        w_r[32 * lo:32 * (lo + cnt)] = d_w[32 * lo:32 * (lo + cnt)]
        h_r = torch.zeros_like(d_h)
        h_lo, h_cnt = ranges[r][4], ranges[r][5]
        h_r[32 * h_lo:32 * (h_lo + h_cnt)] = d_h[32 * h_lo:32 * (h_lo + h_cnt)]
        torch.cuda.synchronize()
        part = d_all[768 * r:]
        bn.ctx.check(bn.lib.ws_bn128_groth16_witness_msms_dev(bn.ctx.h, keys[r], w_r.data_ptr(), part.data_ptr(), None))
        bn.ctx.check(bn.lib.ws_bn128_groth16_h_msm_dev(bn.ctx.h, keys[r], h_r.data_ptr() + 32 * h_lo, part.data_ptr(), None))
        bn.ctx.timings()
        torch.cuda.synchronize()
    bn.ctx.check(bn.lib.ws_bn128_groth16_finish_dev(bn.ctx.h, keys[0], d_all.data_ptr(), world, d_rs.data_ptr(), d_out.data_ptr(), None))
    bn.ctx.timings()
    torch.cuda.synchronize()
    out = bytes(d_out.cpu().numpy())
    assert {"pi_a": bn.bin2g1(out[:96]), "pi_b": bn.bin2g2(out[96:288]), "pi_c": bn.bin2g1(out[288:])} == g["proof_rs"]
    for k_ in keys:
        bn.lib.ws_bn128_pkey_free(k_)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_balanced_split_emulated_on_one_gpu(bn, world):
    """ws_bn128_pkey_load_ranges with parallel.balanced_ranges pieces (a rank holds one or two large contiguous pieces of
    the five sections, some sections not at all), every 'rank' run on this GPU, records concatenated, finish: golden."""
    import ctypes
    import torch
    from wasmsnark_b200._lib import as_ptr
    from wasmsnark_b200.parallel import balanced_ranges, section_weights
    pk, wit, g = _synth_small()
    weights, sizes = section_weights(pk)
    d_w = torch.frombuffer(bytearray(wit), dtype=torch.uint8).cuda()
    d_rs = torch.frombuffer(bytearray(bytes.fromhex(g["r"]) + bytes.fromhex(g["s"])), dtype=torch.uint8).cuda()
    d_all = torch.zeros(768 * world, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(384, dtype=torch.uint8, device="cuda")
    d_h = torch.zeros(256 * 32, dtype=torch.uint8, device="cuda")
    keys = []
    for r in range(world):
        k = ctypes.c_void_p()
        arr = (ctypes.c_uint32 * 10)(*balanced_ranges(weights, sizes, r, world))
        bn.ctx.check(bn.lib.ws_bn128_pkey_load_ranges(bn.ctx.h, as_ptr(pk), len(pk), arr, ctypes.byref(k)))
        keys.append(k)
    bn.ctx.check(bn.lib.ws_bn128_groth16_h_dev(bn.ctx.h, keys[world - 1], d_w.data_ptr(), d_h.data_ptr(), None))
    bn.ctx.timings()
    torch.cuda.synchronize()
    for r in range(world):
        rg = (ctypes.c_uint32 * 6)()
        bn.lib.ws_bn128_pkey_shard_info(keys[r], rg)
        part = d_all[768 * r:]
        bn.ctx.check(bn.lib.ws_bn128_groth16_witness_msms_dev(bn.ctx.h, keys[r], d_w.data_ptr(), part.data_ptr(), None))
        bn.ctx.check(bn.lib.ws_bn128_groth16_h_msm_dev(bn.ctx.h, keys[r], d_h.data_ptr() + 32 * rg[4], part.data_ptr(), None))
        bn.ctx.timings()
        torch.cuda.synchronize()
    bn.ctx.check(bn.lib.ws_bn128_groth16_finish_dev(bn.ctx.h, keys[0], d_all.data_ptr(), world, d_rs.data_ptr(), d_out.data_ptr(), None))
    bn.ctx.timings()
    torch.cuda.synchronize()
    out = bytes(d_out.cpu().numpy())
    assert {"pi_a": bn.bin2g1(out[:96]), "pi_b": bn.bin2g2(out[96:288]), "pi_c": bn.bin2g1(out[288:])} == g["proof_rs"]
    # a partial key is refused by the whole-proof entry point
    from wasmsnark_b200._lib import WsError
    with pytest.raises(WsError):
        bn.ctx.check(bn.lib.ws_bn128_groth16_prove_dev(bn.ctx.h, keys[0], d_w.data_ptr(), d_rs.data_ptr(), d_out.data_ptr(), None))
    for k_ in keys:
        bn.lib.ws_bn128_pkey_free(k_)


def test_back_to_back_async_proofs_without_a_sync_between_them(bn):
    """The pipelined form bench.py times: several ws_bn128_groth16_prove_dev calls queued on ONE context and ONE stream with no
    synchronisation between them -- different witnesses, different (r, s), one output buffer each -- and only then a sync.
    Every call shares the context's workspace (MSM arenas, the H buffers, the side streams), so the library has to order proof
    i + 1 behind proof i ON THE DEVICE; every one of the proofs must be the right one (the bench itself only sees the last).
    Under the emulator's deferred modes (CUEMU_ASYNC) this is where a missing dependency between consecutive proofs shows."""
    import ctypes
    import torch
    pk, wit, g = _synth_small()
    n = len(wit) // 32
    rnd = random.Random(17)
    cases = []
    for i in range(5):
        w = bytearray(wit)
        for k in rnd.sample(range(1, n), 40):
            w[32 * k:32 * k + 32] = rnd.randrange(O.R_ORDER).to_bytes(32, "little")
        cases.append((bytes(w), rnd.getrandbits(256).to_bytes(32, "little"), rnd.getrandbits(256).to_bytes(32, "little")))
    cases[0] = (wit, bytes.fromhex(g["r"]), bytes.fromhex(g["s"]))             # the golden one among them
    key = bn.load_proving_key(pk)
    try:
        d_w = [torch.frombuffer(bytearray(w), dtype=torch.uint8).cuda() for w, _, _ in cases]
        d_rs = [torch.frombuffer(bytearray(r + s), dtype=torch.uint8).cuda() for _, r, s in cases]
        d_out = [torch.zeros(384, dtype=torch.uint8, device="cuda") for _ in cases]
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            for i in range(len(cases)):
                bn.ctx.check(bn.lib.ws_bn128_groth16_prove_dev(bn.ctx.h, key.h, d_w[i].data_ptr(), d_rs[i].data_ptr(), d_out[i].data_ptr(),
                                                               ctypes.c_void_p(st.cuda_stream)))
        if os.environ.get("CUEMU_ASYNC") and hasattr(bn.lib, "cuemu_pending_ops"):        # emulator, deferred: nothing has run yet
            bn.lib.cuemu_pending_ops.restype = ctypes.c_ulonglong
            assert bn.lib.cuemu_pending_ops() > 100 * len(cases), bn.lib.cuemu_pending_ops()
        st.synchronize()
        for i, (w, r, s) in enumerate(cases):
            out = bytes(d_out[i].cpu().numpy())
            got = {"pi_a": bn.bin2g1(out[:96]), "pi_b": bn.bin2g2(out[96:288]), "pi_c": bn.bin2g1(out[288:])}
            assert got == O.groth16_gen_proof(w, pk, int.from_bytes(r, "little"), int.from_bytes(s, "little")), i
        assert O.groth16_gen_proof(*cases[0][:1], pk, int.from_bytes(cases[0][1], "little"), int.from_bytes(cases[0][2], "little")) == g["proof_rs"]
    finally:
        key.free()


def test_context_is_thread_safe(bn):
    """Several host threads proving / multiplying on ONE context (the reference queues any number of commands,
    src/bn128.js:282-300; the Node addon runs proofs on libuv worker threads): every result is the right one."""
    import threading
    pk, wit, g = _synth_small()
    r, s = bytes.fromhex(g["r"]), bytes.fromhex(g["s"])
    key = bn.load_proving_key(pk)
    errors = []

    def prover(n):
        try:
            for _ in range(n):
                assert bn.groth16GenProof(wit, key, r=r, s=s) == g["proof_rs"]
        except Exception as e:       # noqa: BLE001
            errors.append(e)

    def stateless(n):
        try:
            for _ in range(n):
                assert bn.groth16GenProof(wit, pk, r=r, s=s) == g["proof_rs"]
        except Exception as e:       # noqa: BLE001
            errors.append(e)

    def ffts(n):
        try:
            x = bytes(range(32)) * 64
            want = bn.fft(x)
            for _ in range(n):
                assert bn.fft(x) == want
        except Exception as e:       # noqa: BLE001
            errors.append(e)
    th = [threading.Thread(target=f, args=(6,)) for f in (prover, prover, stateless, ffts)]
    [t.start() for t in th]
    [t.join() for t in th]
    key.free()
    assert not errors, errors


def _witness_near_2p256(w):
    """raw signals in the top 1/8 of the 256-bit range (legal for the reference: frm_toMontgomeryN takes them as they come,
    src/bn128.js:134), among them values for which a device product with the raw operand FIRST drops a carry
    (tests/test_fp_asm_model.py): 2^256 - 2, 2^256 - 2^32, ..."""
    rnd = random.Random(9)
    top = (1 << 256) - 1
    vals = [top - 1, top ^ 0xFFFFFFFF, top ^ (0xFFFFFFFF << 224), top - O.R_ORDER + 1]
    n = len(w) // 32
    vals = (vals + [rnd.randrange(7 << 253, 1 << 256) for _ in range(n)])[:n]
    return b"".join(v.to_bytes(32, "little") for v in vals)


@pytest.mark.parametrize("case", ["C empty", "one signal", "domain 1", "domain 2", "nVars < 32", "all infinity", "no infinity", "witness 0",
                                  "witness 2^256-1", "witness near 2^256", "no coefficients", "dense coefficients"])
def test_degenerate_keys_and_witnesses(bn, case):
    """Corners the example key never reaches: an empty C section (nVars = nPublic + 1), one-signal keys, one- and two-point
    domains (transforms of size 1, 2, 4), sections that are all infinity or never infinity, witnesses of all zeros / all
    2^256 - 1 (unreduced), polynomial sections without any or with many coefficients -- stateless and resident-key proofs
    against the big-integer restatement of src/bn128.js:580-720 (oracle/bn128.py, itself pinned on the reference's proofs)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import gen_synth_key
    nv, npub, dom, kw, wit_of = {
        "C empty": (9, 8, 8, {}, None),
        "one signal": (1, 0, 1, {}, None),
        "domain 1": (2, 1, 1, {}, None),
        "domain 2": (5, 1, 2, {}, None),
        "nVars < 32": (3, 1, 4, {}, None),
        "all infinity": (200, 5, 128, {"inf_share": 1.0}, None),
        "no infinity": (200, 5, 128, {"inf_share": 0.0}, None),
        "witness 0": (200, 5, 128, {}, lambda w: bytes(len(w))),
        "witness 2^256-1": (200, 5, 128, {}, lambda w: b"\xff" * len(w)),
        "witness near 2^256": (200, 5, 128, {}, _witness_near_2p256),
        "no coefficients": (100, 5, 64, {"coef_density": (0.0, 0.0)}, None),
        "dense coefficients": (100, 5, 64, {"coef_density": (6.0, 9.0)}, None),
    }[case]
    pk, wit = gen_synth_key.make(nv, npub, dom, seed=3, **kw)
    if wit_of:
        wit = wit_of(wit)
    r, s = bytes(range(7, 39)), bytes(range(100, 132))
    want = O.groth16_gen_proof(wit, pk, int.from_bytes(r, "little"), int.from_bytes(s, "little"))
    assert bn.groth16GenProof(wit, pk, r=r, s=s) == want
    key = bn.load_proving_key(pk)
    try:
        assert bn.groth16GenProof(wit, key, r=r, s=s) == want
    finally:
        key.free()


def int_edge_values(n=8, bits=32):
    """helpers.genValues(n, false) of the reference's tests (test/helpers/helpers.js:3-23): 0, 2^(32 i) - 1, 2^(32 i), 2^(32 i) + 1"""
    res = [0]
    for i in range(n):
        if i > 0:
            res.append((1 << (bits * i)) - 1)
        if i < n - 1:
            res += [1 << (bits * i), (1 << (bits * i)) + 1]
    return res


def test_int_mul_edge_grid_as_r_times_s(bn):
    """test/int.js:22-47 multiplies every pair of its edge patterns with int_mul and compares the 512-bit products.  On this
    path int_mul has one job: r * s, exact, as the 512-bit scalar of (r s) delta1 in pi_c (src/bn128.js:701) -- so every pair
    of the grid, plus 2^256 - 1, goes in as (r, s) of a proof (tiny resident key), against the big-integer prover; r and s
    alone also drive the fixed-base products r delta1, s delta1, s delta2 and the windowed s pi_a, r pib1."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import gen_synth_key
    pk, wit = gen_synth_key.make(6, 1, 4, seed=4)
    vals = int_edge_values() + [(1 << 256) - 1]
    if os.environ.get("WS_TEST_EMULATED"):          # the CPU emulator takes a second per proof: a diagonal band of the grid
        pairs = [(a, vals[(i + d) % len(vals)]) for i, a in enumerate(vals) for d in (0, 1, 11)]
    else:
        pairs = [(a, b) for a in vals for b in vals]
    key = bn.load_proving_key(pk)
    try:
        for a, b in pairs:
            got = bn.groth16GenProof(wit, key, r=a.to_bytes(32, "little"), s=b.to_bytes(32, "little"))
            assert got == O.groth16_gen_proof(wit, pk, a, b), (hex(a), hex(b))
    finally:
        key.free()


def test_last_error_text_belongs_to_the_failing_thread(bn):
    """Two host threads fail differently on ONE context, and only then ask for the text: each reads its own message
    (ws_last_error keeps a per-thread copy; with a single per-context string the later failure would win for both --
    the Node addon reads the text on a libuv worker thread right after its blocking call returned)."""
    import ctypes
    import threading
    lib, h = bn.lib, bn.ctx.h
    both_failed = threading.Barrier(2)
    got = {}

    def wrong_size():
        buf = ctypes.create_string_buffer(96)
        rc = lib.ws_bn128_fr_fft(h, ctypes.cast(buf, ctypes.c_void_p), 3, 0)
        both_failed.wait(timeout=60)
        got["size"] = (rc, lib.ws_last_error(h).decode())

    def null_buffer():
        rc = lib.ws_bn128_fr_ifft(h, None, 4, 0)
        both_failed.wait(timeout=60)
        got["null"] = (rc, lib.ws_last_error(h).decode())
    th = [threading.Thread(target=f) for f in (wrong_size, null_buffer)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert got["size"][0] != 0 and "power of two" in got["size"][1], got
    assert got["null"][0] != 0 and "null buffer" in got["null"][1], got


def test_stateless_call_from_pageable_memory_large_witness(bn):
    """The pinned staging ring (host_io.cuh) is only used above 1 MiB: a 2^15-signal witness from ordinary bytes objects
    through the stateless call, twice (second call: cached key, fingerprint overlapped), equals the resident-key proof."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import gen_synth_key
    nv = 1 << 15
    pk, wit = gen_synth_key.make(nv + 1000, 7, nv, seed=5)
    assert len(wit) > (1 << 20)
    r, s = bytes(range(32)), bytes(range(32, 64))
    key = bn.load_proving_key(pk)
    want = bn.groth16GenProof(wit, key, r=r, s=s)
    key.free()
    for _ in range(3):
        assert bn.groth16GenProof(bytes(wit), bytes(pk), r=r, s=s) == want
    assert bn.groth16GenProof(bytearray(wit), bytearray(pk), r=r, s=s) == want


def test_multi_handle_on_the_visible_gpus():
    """ws_multi_*: the one-process multi-GPU handle (all visible GPUs; 1 on a single-GPU box, where the code path is the
    same with n = 1).  Golden proofs, MSMs against the single-context results, stateless key replacement."""
    import torch
    from wasmsnark_b200 import build_bn128, build_bn128_multi
    import util
    pk, wit, g = _synth_small()
    r, s = bytes.fromhex(g["r"]), bytes.fromhex(g["s"])
    m = build_bn128_multi(0)
    bn1 = build_bn128(0)
    try:
        assert m.ngpus == torch.cuda.device_count()
        for _ in range(3):          # H owner rotates over the GPUs
            assert m.groth16GenProof(wit, pk, r=r, s=s) == g["proof_rs"]
        assert m.groth16GenProof(wit, pk, r=bytes(32), s=bytes(32)) == g["proof_r0s0"]
        for mode in (1, 0):                             # balanced and uniform point partitions
            m.lib.ws_multi_set_split(m.h, mode)
            pk_mod = pk[:-64] + pk[-128:-64]            # another key (last H point duplicated): forces a reload with this split
            assert m.groth16GenProof(wit, pk_mod, r=r, s=s) == bn1.groth16GenProof(wit, pk_mod, r=r, s=s)
            assert m.groth16GenProof(wit, pk, r=r, s=s) == g["proof_rs"]
        tok = m.load_proving_key(pk)                    # resident form: no per-call key fingerprint
        for _ in range(2):
            assert m.groth16GenProof(wit, tok, r=r, s=s) == g["proof_rs"]
        k = O.parse_pkey(pk)
        a = k["pPointsA"]
        pk3 = pk[:a + 64] + pk[a + 128:a + 192] + pk[a + 64:a + 128] + pk[a + 192:]      # same length, other content
        assert m.groth16GenProof(wit, pk3, r=r, s=s) == bn1.groth16GenProof(wit, pk3, r=r, s=s) != g["proof_rs"]
        assert m.groth16GenProof(wit, pk, r=r, s=s) == g["proof_rs"]
        n = 5000
        aff, P0, D = util.arith_points(O.Fq, O.G1_GEN, n, 11)
        rnd = random.Random(12)
        sc = [rnd.getrandbits(256) for _ in range(n)]
        sb, pb = b"".join(util.b32(v) for v in sc), util.g1_points_bytes(aff)
        assert util.g1_result_tuple(bn1.g1_affine(m.g1_multiexp(sb, pb))) == util.closed_form_sum(sc, P0, D)
        g2p = b"".join(util.g2_affine_m(O.pt_times_scalar(O.Fq2, (O.G2_GEN[0], O.G2_GEN[1], O.Fq2.one), rnd.randrange(1, O.R_ORDER))) for _ in range(40))
        sb2 = b"".join(util.b32(rnd.getrandbits(256)) for _ in range(40))
        assert bn1.g2_affine(m.g2_multiexp(sb2, g2p)) == bn1.g2_affine(bn1.g2_multiexp(sb2, g2p))
    finally:
        m.terminate()
        bn1.terminate()


def test_multi_handle_with_more_gpus_than_points():
    """Keys so small that some GPUs hold EMPTY ranges of every section (nVars 5, domain 2 over up to 8 devices; uniform and
    balanced split), every device taking its turn as H owner, and multiexps of 0 and 3 pairs: the empty shards contribute
    the point at infinity.  On a one-GPU box this is the n = 1 path; the CPU suite runs it over eight pretend devices
    (tests/test_emulated_library.py)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import gen_synth_key
    import util
    from wasmsnark_b200 import build_bn128, build_bn128_multi
    r, s = bytes(range(7, 39)), bytes(range(100, 132))
    ri, si = int.from_bytes(r, "little"), int.from_bytes(s, "little")
    m, bn1 = build_bn128_multi(0), build_bn128(0)
    try:
        for mode, (nv, npub, dom, seed) in ((0, (5, 1, 2, 3)), (1, (6, 1, 4, 4))):
            m.lib.ws_multi_set_split(m.h, mode)
            pk, wit = gen_synth_key.make(nv, npub, dom, seed=seed)
            want = O.groth16_gen_proof(wit, pk, ri, si)
            for _ in range(m.ngpus + 1):
                assert m.groth16GenProof(wit, pk, r=r, s=s) == want
        aff, P0, D = util.arith_points(O.Fq, O.G1_GEN, 3, 9)
        sc = [5, (1 << 255) + 7, 11]
        got = bn1.g1_affine(m.g1_multiexp(b"".join(util.b32(v) for v in sc), util.g1_points_bytes(aff)))
        assert util.g1_result_tuple(got) == util.closed_form_sum(sc, P0, D)
        assert util.g1_result_tuple(bn1.g1_affine(m.g1_multiexp(b"", b"")))[2] == 0
    finally:
        m.terminate()
        bn1.terminate()


def test_golden_proof_2p20_constraints(bn):
    """BASELINE's headline config: the 2^20-constraint synthetic key (tools/gen_synth_key.make(2^20, 58, 2^20, seed 0xB204))
    with the bench's (r, s) against the proof the REFERENCE's own code computed for it (tests/golden/make_golden_2p20.py,
    pkey sha256 pinned).  Key generation dominates (~20 s of host time)."""
    import hashlib
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import gen_synth_key
    gold = json.load(open(os.path.join(GOLD, "synth_2p20.golden.json")))
    k = gold["log2_constraints"]
    pk, wit = gen_synth_key.make(1 << k, gold["n_public"], 1 << k, seed=gold["seed"])
    assert hashlib.sha256(pk).hexdigest() == gold["pkey_sha256"]
    assert hashlib.sha256(wit).hexdigest() == gold["witness_sha256"]
    key = bn.load_proving_key(pk)
    try:
        proof = bn.groth16GenProof(wit, key, r=bytes.fromhex(gold["r"]), s=bytes.fromhex(gold["s"]))
    finally:
        key.free()
    assert proof == gold["proof"]
