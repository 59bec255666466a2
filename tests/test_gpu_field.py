"""C1 / C5 field parity on the GPU, through the C ABI: 1k random pairs + the reference's edge
grid (test/f1.js:299-344) for BN128 Fq, Fr and MNT6753 Fq, bit-exact against the oracle and,
for mul, against the reference's own f1m_mul / frm_mul (compiled WASM image)."""
import random

import numpy as np
import pytest

from oracle import bn128 as O
from oracle import ref_wasm
from wasmsnark_b200 import _lib
from util import MNT_Q, Q, RR, edge_values

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bn():
    from wasmsnark_b200 import build_bn128
    b = build_bn128(0)
    yield b
    b.terminate()


def pairs(p, seed):
    rng = np.random.Generator(np.random.PCG64(seed))        # PCG64(0xB200) 256-bit draws mod p
    nb = (p.bit_length() + 7) // 8 + 8
    draw = lambda: int.from_bytes(rng.bytes(nb), "little") % p
    xs = [draw() for _ in range(1000)]
    ys = [draw() for _ in range(1000)]
    ev = edge_values(p)
    for a in ev:
        for c in ev[:6]:
            xs.append(a)
            ys.append(c)
    return xs, ys


@pytest.mark.parametrize("field,p,esz", [(_lib.FIELD_BN128_FQ, Q, 32), (_lib.FIELD_BN128_FR, RR, 32), (_lib.FIELD_MNT6753_FQ, MNT_Q, 96)])
def test_field_ops_bit_exact(bn, field, p, esz):
    xs, ys = pairs(p, 0xB200)
    R = 1 << (8 * esz)
    Rinv = pow(R, -1, p)
    X = b"".join(v.to_bytes(esz, "little") for v in xs)
    Y = b"".join(v.to_bytes(esz, "little") for v in ys)
    n = len(xs)

    def unpack(bs):
        return [int.from_bytes(bs[i * esz:(i + 1) * esz], "little") for i in range(n)]

    assert unpack(bn.fp_batch(field, _lib.OP_MUL, X, Y)) == [x * y * Rinv % p for x, y in zip(xs, ys)]
    assert unpack(bn.fp_batch(field, _lib.OP_ADD, X, Y)) == [(x + y) % p for x, y in zip(xs, ys)]
    assert unpack(bn.fp_batch(field, _lib.OP_SUB, X, Y)) == [(x - y) % p for x, y in zip(xs, ys)]
    assert unpack(bn.fp_batch(field, _lib.OP_SQUARE, X)) == [x * x * Rinv % p for x in xs]
    assert unpack(bn.fp_batch(field, _lib.OP_TOMONT, X)) == [x * R % p for x in xs]
    assert unpack(bn.fp_batch(field, _lib.OP_FROMMONT, X)) == [x * Rinv % p for x in xs]
    m = 64
    got = unpack(bn.fp_batch(field, _lib.OP_INVERSE, X))[:m]
    assert got == [(pow(x * Rinv % p, -1, p) * R % p if x else 0) for x in xs[:m]]


def raw_values(p, esz, seed):
    """numbers that fit the element's bytes but need not be below p, dense at the top of the range"""
    rnd = random.Random(seed)
    top = (1 << (8 * esz)) - 1
    lim = top - 2 * p if top > 4 * p else p
    return [0, 1, p - 1, p, p + 1, top, top - 1, top ^ 0xFFFFFFFF, top ^ (0xFFFFFFFF << (8 * esz - 32)), top - p, top - p + 1, 1 << (8 * esz - 1)] + \
           [rnd.randrange(lim, top + 1) for _ in range(700)] + [rnd.randrange(top + 1) for _ in range(300)]


@pytest.mark.parametrize("field,p,esz", [(_lib.FIELD_BN128_FQ, Q, 32), (_lib.FIELD_BN128_FR, RR, 32), (_lib.FIELD_MNT6753_FQ, MNT_Q, 96)])
def test_to_montgomery_takes_raw_values(bn, field, p, esz):
    """f1m_toMontgomery / frm_toMontgomery accept ANY number that fits the element (the witness enters the prover that way,
    src/bn128.js:134) and return the canonical x R mod p -- checked against the definition and, for BN128, against the
    reference's own functions on the same bytes.  (The device product needs the raw value as its SECOND operand:
    tests/test_fp_asm_model.py, tests/test_ptx_kernels.py.)"""
    xs = raw_values(p, esz, 21)
    X = b"".join(v.to_bytes(esz, "little") for v in xs)
    got = bn.fp_batch(field, _lib.OP_TOMONT, X)
    assert [int.from_bytes(got[i * esz:(i + 1) * esz], "little") for i in range(len(xs))] == [(x << (8 * esz)) % p for x in xs]
    if esz == 32 and ref_wasm.available():
        inst = ref_wasm.RefInstance(pages=100)
        px, pr = inst.alloc(32), inst.alloc(32)
        fn = "f1m_toMontgomery" if field == _lib.FIELD_BN128_FQ else "frm_toMontgomery"
        for i in range(len(xs)):
            inst.put(px, X[32 * i:32 * i + 32])
            inst.call(fn, px, pr)
            assert inst.get(pr, 32) == got[32 * i:32 * i + 32], (fn, hex(xs[i]))
        inst.close()


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not built")
def test_mul_against_reference_image(bn):
    """config C1: 1k random pairs vs the reference's f1m_mul / frm_mul, byte for byte."""
    inst = ref_wasm.RefInstance(pages=1000)
    px, py, pr = inst.alloc(32), inst.alloc(32), inst.alloc(32)
    for field, p, fn in ((_lib.FIELD_BN128_FQ, Q, "f1m_mul"), (_lib.FIELD_BN128_FR, RR, "frm_mul")):
        xs, ys = pairs(p, 0xB200)
        X = b"".join(v.to_bytes(32, "little") for v in xs)
        Y = b"".join(v.to_bytes(32, "little") for v in ys)
        got = bn.fp_batch(field, _lib.OP_MUL, X, Y)
        for i in range(len(xs)):
            inst.put(px, X[32 * i:32 * i + 32])
            inst.put(py, Y[32 * i:32 * i + 32])
            inst.call(fn, px, py, pr)
            assert inst.get(pr, 32) == got[32 * i:32 * i + 32], (fn, i)
    inst.close()


def test_fq2_mul(bn):
    rnd = random.Random(5)
    vals = [(0, 0), (1, 0), (0, 1), (Q - 1, Q - 1)] + [(rnd.randrange(Q), rnd.randrange(Q)) for _ in range(500)]
    xs = vals
    ys = vals[::-1]
    enc = lambda v: O.to_mont(v[0]).to_bytes(32, "little") + O.to_mont(v[1]).to_bytes(32, "little")
    got = bn.fq2_mul_batch(b"".join(enc(v) for v in xs), b"".join(enc(v) for v in ys))
    for i, (x, y) in enumerate(zip(xs, ys)):
        g = (O.from_mont(int.from_bytes(got[64 * i:64 * i + 32], "little")), O.from_mont(int.from_bytes(got[64 * i + 32:64 * i + 64], "little")))
        assert g == O.Fq2.mul(x, y)
