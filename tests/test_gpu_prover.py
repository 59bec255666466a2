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
