"""The PTX nvcc emits for the library's element-wise field kernels, EXECUTED on the CPU (tests/ptx_exec.py) against the
definitions -- the one layer the kernel emulator cannot reach (there the field arithmetic takes its portable host path).

  k_fr_to_mont (prover.cu)       how the witness enters the H pipeline (frm_toMontgomeryN, src/bn128.js:134): raw 256-bit values
  k_fp_batch<field, op> (api.cu)  f1m_mul / add / sub / square / toMontgomery / fromMontgomery (build_f1m.js) for BN128 q and r

CPU only; needs nvcc (the build needs it anyway).  ptxas' translation of this PTX is what the B200 run checks."""
import random

import pytest

import ptx_exec as X

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
RR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
TOP = (1 << 256) - 1


def raw_values(p, rnd, n):
    edge = [0, 1, p - 1, p, p + 1, TOP, TOP - 1, TOP ^ 0xFFFFFFFF, TOP ^ (0xFFFFFFFF << 224), TOP - p, TOP - p + 1, 4 * p, 5 * p, 5 * p + 1, 1 << 255]
    return edge + [rnd.randrange(4 * p, 1 << 256) for _ in range((n - len(edge)) * 2 // 3)] + [rnd.randrange(1 << 256) for _ in range(n - len(edge) - (n - len(edge)) * 2 // 3)]


def b32(vals):
    return b"".join(v.to_bytes(32, "little") for v in vals)


def ints(b):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def test_witness_to_montgomery_kernel_on_raw_256_bit_values():
    """k_fr_to_mont as compiled: canonical a * 2^256 mod r for every raw a, the top of the range included (the operand order
    of fp.cuh::to_mont matters there: tests/test_fp_asm_model.py) -- 300 values over three blocks, the guard `i < n` included"""
    params, stmts = X.entry(X.ptx_of("prover.cu"), "k_fr_to_mont")
    vals = raw_values(RR, random.Random(11), 300)
    n = len(vals)
    mem = X.Memory()
    src, dst = mem.alloc(b32(vals)), mem.alloc(b"\xAA" * (32 * (n + 4)))
    X.launch(params, stmts, [src, dst, n], mem, (n + 127) // 128, 128)
    out = ints(mem.bytes_at(dst))
    assert out[:n] == [(a << 256) % RR for a in vals]
    assert out[n:] == [int.from_bytes(b"\xAA" * 32, "little")] * 4           # nothing written beyond n


@pytest.mark.parametrize("field,p", [("BN128_Fq", Q), ("BN128_Fr", RR)])
@pytest.mark.parametrize("op", ["mul", "add", "sub", "square", "toMontgomery", "fromMontgomery"])
def test_field_batch_kernels_as_compiled(field, p, op):
    """k_fp_batch<field, op> as compiled, on the edge values of the reference's field tests (test/f1.js) and random pairs;
    toMontgomery on raw values"""
    code = ["mul", "add", "sub", "square", "toMontgomery", "fromMontgomery"].index(op)
    R = 1 << 256
    rinv = pow(R, -1, p)
    want = {"mul": lambda a, b: a * b * rinv % p, "add": lambda a, b: (a + b) % p, "sub": lambda a, b: (a - b) % p, "square": lambda a, b: a * a * rinv % p,
            "toMontgomery": lambda a, b: a * R % p, "fromMontgomery": lambda a, b: a * rinv % p}[op]
    params, stmts = X.entry(X.ptx_of("api.cu"), ["k_fp_batch", field, "ELi%dE" % code])
    rnd = random.Random(code * 7 + p % 13)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, 1 << 253, (1 << 253) - 1, 0xFFFFFFFF, 1 << 32, (1 << 224) - 1]
    xs = [x for x in edge for _ in edge] + [rnd.randrange(p) for _ in range(48)]
    ys = [y for _ in edge for y in edge] + [rnd.randrange(p) for _ in range(48)]
    if op == "toMontgomery":
        xs = raw_values(p, rnd, len(xs))
    n = len(xs)
    mem = X.Memory()
    dx, dy, do = mem.alloc(b32(xs)), mem.alloc(b32(ys)), mem.alloc(bytes(32 * n))
    X.launch(params, stmts, [dx, dy, do, n], mem, (n + 127) // 128, 128)
    assert ints(mem.bytes_at(do)) == [want(a, b) for a, b in zip(xs, ys)]


def test_executor_refuses_what_it_does_not_model():
    """a kernel outside the integer subset (here: the inversion, which keeps a table in local memory) fails loudly"""
    params, stmts = X.entry(X.ptx_of("api.cu"), ["k_fp_batch", "BN128_Fq", "ELi6E"])
    mem = X.Memory()
    d = mem.alloc(bytes(32))
    with pytest.raises((NotImplementedError, ValueError, KeyError)):
        X.launch(params, stmts, [d, d, d, 1], mem, 1, 1)
