"""The PTX nvcc emits for the library's element-wise field kernels, EXECUTED on the CPU (tests/ptx_exec.py) against the
definitions -- the one layer the kernel emulator cannot reach (there the field arithmetic takes its portable host path).

  k_fr_to_mont (prover.cu)       how the witness enters the H pipeline (frm_toMontgomeryN, src/bn128.js:134): raw 256-bit values
  k_fp_batch<field, op> (api.cu, bls12381_g1.cu)  f1m_mul / add / sub / square / toMontgomery / fromMontgomery (build_f1m.js)
                                 for BN128 q and r (8 limbs), BLS12-381 q (12 limbs) and r (8 limbs, 255 bits: no dedicated
                                 squaring) and MNT6753 q (24 limbs: the product is an out-of-line device function reached
                                 through local memory); f1m_inverse (the Fermat ladder: loops, a local table) for BN128

CPU only; needs nvcc (the build needs it anyway).  ptxas' translation of this PTX is what the B200 run checks."""
import random

import pytest

import ptx_exec as X
from util import MNT_Q

@pytest.fixture(scope="module")
def ptx():
    return X.ptx_of_many(["prover.cu", "api.cu", "bls12381_g1.cu"])


Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
BLS_Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
BLS_R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
RR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
TOP = (1 << 256) - 1


def raw_values(p, rnd, n, esz=32):
    """numbers that fit the element's bytes, dense at the top of the range"""
    top = (1 << (8 * esz)) - 1
    lo = top - 2 * p if top > 4 * p else p
    edge = [0, 1, p - 1, p, p + 1, top, top - 1, top ^ 0xFFFFFFFF, top ^ (0xFFFFFFFF << (8 * esz - 32)), top - p, top - p + 1, 2 * p, 2 * p + 1, 1 << (8 * esz - 1)]
    return edge + [rnd.randrange(lo, top + 1) for _ in range((n - len(edge)) * 2 // 3)] + [rnd.randrange(top + 1) for _ in range(n - len(edge) - (n - len(edge)) * 2 // 3)]


def b32(vals, esz=32):
    return b"".join(v.to_bytes(esz, "little") for v in vals)


def ints(b, esz=32):
    return [int.from_bytes(b[i:i + esz], "little") for i in range(0, len(b), esz)]


def test_witness_to_montgomery_kernel_on_raw_256_bit_values(ptx):
    """k_fr_to_mont as compiled: canonical a * 2^256 mod r for every raw a, the top of the range included (the operand order
    of fp.cuh::to_mont matters there: tests/test_fp_asm_model.py) -- 300 values over three blocks, the guard `i < n` included"""
    params, stmts = X.entry(ptx["prover.cu"], "k_fr_to_mont")
    vals = raw_values(RR, random.Random(11), 300)
    n = len(vals)
    mem = X.Memory()
    src, dst = mem.alloc(b32(vals)), mem.alloc(b"\xAA" * (32 * (n + 4)))
    X.launch(params, stmts, [src, dst, n], mem, (n + 127) // 128, 128)
    out = ints(mem.bytes_at(dst))
    assert out[:n] == [(a << 256) % RR for a in vals]
    assert out[n:] == [int.from_bytes(b"\xAA" * 32, "little")] * 4           # nothing written beyond n


@pytest.mark.parametrize("field,p,esz,unit", ids=lambda v: v if isinstance(v, str) else "", argvalues=[("BN128_Fq", Q, 32, "api.cu"), ("BN128_Fr", RR, 32, "api.cu"),
                                              ("BLS12_381_Fq", BLS_Q, 48, "bls12381_g1.cu"), ("BLS12_381_Fr", BLS_R, 32, "bls12381_g1.cu"),
                                              ("MNT6753_Fq", MNT_Q, 96, "api.cu")])
@pytest.mark.parametrize("op", ["mul", "add", "sub", "square", "toMontgomery", "fromMontgomery"])
def test_field_batch_kernels_as_compiled(ptx, field, p, esz, unit, op):
    """k_fp_batch<field, op> as compiled, on the edge values of the reference's field tests (test/f1.js) and random pairs;
    toMontgomery on raw values"""
    code = ["mul", "add", "sub", "square", "toMontgomery", "fromMontgomery"].index(op)
    R = 1 << (8 * esz)
    rinv = pow(R, -1, p)
    want = {"mul": lambda a, b: a * b * rinv % p, "add": lambda a, b: (a + b) % p, "sub": lambda a, b: (a - b) % p, "square": lambda a, b: a * a * rinv % p,
            "toMontgomery": lambda a, b: a * R % p, "fromMontgomery": lambda a, b: a * rinv % p}[op]
    params, stmts = X.entry(ptx[unit], ["k_fp_batch", field, "ELi%dE" % code])
    rnd = random.Random(code * 7 + p % 13)
    hb = p.bit_length() - 1
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, 1 << hb, (1 << hb) - 1, 0xFFFFFFFF, 1 << 32, (1 << (8 * esz - 32)) - 1]
    if esz == 96:                                  # 24 limbs: 1300 instructions per product and thread -- a smaller sample
        edge = edge[:6]
    xs = [x for x in edge for _ in edge] + [rnd.randrange(p) for _ in range(48 if esz < 96 else 12)]
    ys = [y for _ in edge for y in edge] + [rnd.randrange(p) for _ in range(48 if esz < 96 else 12)]
    if op == "toMontgomery":
        xs = raw_values(p, rnd, len(xs), esz)
    n = len(xs)
    mem = X.Memory()
    dx, dy, do = mem.alloc(b32(xs, esz)), mem.alloc(b32(ys, esz)), mem.alloc(bytes(esz * n))
    X.launch(params, stmts, [dx, dy, do, n], mem, (n + 127) // 128, 128)
    assert ints(mem.bytes_at(do), esz) == [want(a, b) for a, b in zip(xs, ys)]


@pytest.mark.parametrize("field,p", [("BN128_Fq", Q), ("BN128_Fr", RR)])
def test_inversion_kernel_as_compiled(ptx, field, p):
    """k_fp_batch<field, inverse>: a^(p-2) by square-and-multiply (fp.cuh::inv) -- 254 dedicated squarings and the products
    in between, loops and branches as the compiler laid them out; inverse(0) = 0 like f1m_inverse"""
    params, code = X.entry(ptx["api.cu"], ["k_fp_batch", field, "ELi6E"])
    rnd = random.Random(p % 97)
    xs = [0, 1, p - 1] + [rnd.randrange(p) for _ in range(2)]                  # 100k interpreted statements per value
    mem = X.Memory()
    dx, do = mem.alloc(b32(xs)), mem.alloc(bytes(32 * len(xs)))
    X.launch(params, code, [dx, dx, do, len(xs)], mem, 1, 32)
    R = 1 << 256
    assert ints(mem.bytes_at(do)) == [(R * R * pow(x, -1, p)) % p if x else 0 for x in xs]


def test_executor_refuses_what_it_does_not_model(ptx):
    """a kernel outside the subset (here: an NTT pass -- shared memory, barriers) fails loudly instead of being half-executed"""
    import re
    name = next(m for m in re.findall(r"\.entry\s+(\w+)", ptx["prover.cu"]) if "k_ntt_pass" in m)
    params, code = X.entry(ptx["prover.cu"], name)
    mem = X.Memory()
    d = mem.alloc(bytes(1 << 16))
    with pytest.raises((NotImplementedError, AssertionError, KeyError, ValueError)):
        X.launch(params, code, [d] * len(params), mem, 1, 32)
