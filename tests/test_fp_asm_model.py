"""Executes the PTX instruction lists that tools/gen_fp_asm.py emits for the dedicated squaring (the
shift_mad_from / cmad_from variants) in a small Python model of the carry flag, composed exactly as
fp.cuh::sqr_inline composes them, and checks the result against a^2 R^-1 mod p.  CPU only: catches generator /
composition mistakes before any GPU time is spent (tests/test_gpu_field.py re-checks the real kernels)."""
import os
import random
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_fp_asm as G  # noqa: E402

M32 = 0xFFFFFFFF
N = 8
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
RR = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def run(ins, cells):
    """cells: list of [value] boxes in operand order.  Implements the handful of opcodes the generator uses."""
    cc = 0
    for st in ins:
        st = st.strip().rstrip(";")
        op, rest = st.split(" ", 1)
        args = [a.strip() for a in rest.split(",")]

        def val(a):
            return cells[int(a[1:])][0] if a.startswith("%") else int(a)
        d = cells[int(args[0][1:])]
        if op in ("mul.lo.u32", "mul.hi.u32"):
            pr = val(args[1]) * val(args[2])
            d[0] = (pr & M32) if ".lo" in op else (pr >> 32)
            continue
        m = re.match(r"(mad|madc)\.(lo|hi)(\.cc)?\.u32", op)
        if m:
            pr = val(args[1]) * val(args[2])
            t = ((pr & M32) if m.group(2) == "lo" else (pr >> 32)) + val(args[3]) + (cc if m.group(1) == "madc" else 0)
            d[0] = t & M32
            if m.group(3):
                cc = t >> 32
            continue
        m = re.match(r"(add|addc)(\.cc)?\.u32", op)
        assert m, op
        t = val(args[1]) + val(args[2]) + (cc if m.group(1) == "addc" else 0)
        d[0] = t & M32
        if m.group(2):
            cc = t >> 32
    return cc


def limbs(x):
    return [(x >> (32 * i)) & M32 for i in range(N)]


def value(v):
    return sum(x << (32 * i) for i, x in enumerate(v))


# ---- the primitives of fp_asm.cuh as Python functions over lists (in place), from the SAME instruction lists
def shift_mad_from(K, E, O, a, bi):       # a: list pre-offset like the C++ pointer
    H = N // 2
    cells = [[O[i]] for i in range(N)] + [[E[0]]] + [[a[2 * k]] for k in range(K, H)] + [[bi]]
    run(G.shift_mad_from_ins(N, K), cells)
    for i in range(N):
        O[i] = cells[i][0]
    E[0] = cells[N][0]


def cmad_from(K, acc, a, bi):
    H = N // 2
    cells = [[acc[i]] for i in range(N)] + [[0]] + [[a[2 * k]] for k in range(K, H)] + [[bi]]
    run(G.cmad_from_ins(N, K), cells)
    for i in range(N):
        acc[i] = cells[i][0]
    return cells[N][0]


def mul_n(a, bi):
    acc = [0] * N
    for k in range(N // 2):
        pr = a[2 * k] * bi
        acc[2 * k], acc[2 * k + 1] = pr & M32, pr >> 32
    return acc


def merge(e, o):
    c = 0
    for k in range(N - 1):
        t = e[k] + o[k + 1] + c
        e[k], c = t & M32, t >> 32
    e[N - 1] = (e[N - 1] + c) & M32


def sqr_model(a_int, p):
    """mirror of fp.cuh::sqr_inline (N = 8)"""
    np0 = (-pow(p, -1, 1 << 32)) & M32
    a, pl = limbs(a_int), limbs(p)
    d = [0] + [((a[j] << 1) | (a[j - 1] >> 31)) & M32 for j in range(1, N)]

    def xrow(i):
        return [0 if j < i else (a[i] if j == i else ((d[j] & ~1 & M32) if j == i + 1 else d[j])) for j in range(N)] + [0]
    x = xrow(0)
    e, o = mul_n(x, a[0]), mul_n(x[1:], a[0])
    m = (e[0] * np0) & M32
    cmad_from(0, o, pl[1:] + [0], m)
    o[N - 1] = (o[N - 1] + cmad_from(0, e, pl, m)) & M32
    for i in range(1, N):
        E, O = (o, e) if i & 1 else (e, o)
        x = xrow(i)
        shift_mad_from(i >> 1, E, O, x[1:], a[i])
        if ((i + 1) >> 1) < N // 2:
            O[N - 1] = (O[N - 1] + cmad_from((i + 1) >> 1, E, x, a[i])) & M32
        m = (E[0] * np0) & M32
        cmad_from(0, O, pl[1:] + [0], m)
        O[N - 1] = (O[N - 1] + cmad_from(0, E, pl, m)) & M32
    merge(e, o)
    r = value(e)
    return r - p if r >= p else r


def test_k0_variants_are_the_plain_primitives():
    rnd = random.Random(1)
    for _ in range(50):
        E, O = [rnd.getrandbits(32) for _ in range(N)], [rnd.getrandbits(32) for _ in range(N)]
        O[N - 1] = rnd.randrange(4)
        a, bi = [rnd.getrandbits(32) for _ in range(N)], rnd.getrandbits(32)
        E2, O2 = list(E), list(O)
        shift_mad_from(0, E2, O2, a, bi)
        # reference semantics: E0 += O1 (carry c); O' = (O >> 64) + sum a[2k] bi 2^(64k) + c, as a 256-bit value
        t = E[0] + O[1]
        want = (value(O) >> 64) + sum(a[2 * k] * bi << (64 * k) for k in range(N // 2)) + (t >> 32)
        assert E2[0] == t & M32 and value(O2) == want % (1 << 256)


def test_squaring_model_matches_definition():
    rnd = random.Random(2)
    for p in (Q, RR):
        rinv = pow(1 << 256, -1, p)
        cases = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (1 << 253) - 1, (1 << 253), 0xFFFFFFFF, 1 << 32, (1 << 224) - 1] + [rnd.randrange(p) for _ in range(400)]
        for a in cases:
            assert sqr_model(a, p) == a * a * rinv % p, hex(a)


def test_255_bit_modulus_would_overflow_so_it_is_not_enabled():
    """fp.cuh enables the dedicated squaring for p < 2^254 only; the model shows why (BLS12-381 Fr, 255 bits)."""
    p = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    rinv = pow(1 << 256, -1, p)
    rnd = random.Random(3)
    assert any(sqr_model(a, p) != a * a * rinv % p for a in [rnd.randrange(p >> 1, p) for _ in range(200)])
    src = open(os.path.join(ROOT, "wasmsnark_b200", "csrc", "fp.cuh")).read()
    assert "P::BITS <= 254) return sqr_inline8(a)" in src


def mul_model(a_int, b_int, p):
    """mirror of the DEVICE path of fp.cuh::mul_inline (N = 8): WS_ROW0, seven WS_ROWs with the accumulators swapping roles,
    merge_n8, sub_n8 + select on the borrow -- 256-bit truncation included, exactly what the registers hold"""
    np0 = (-pow(p, -1, 1 << 32)) & M32
    a, b, pl = limbs(a_int) + [0], limbs(b_int), limbs(p) + [0]
    e, o = mul_n(a, b[0]), mul_n(a[1:], b[0])
    m = (e[0] * np0) & M32
    cmad_from(0, o, pl[1:], m)
    o[N - 1] = (o[N - 1] + cmad_from(0, e, pl, m)) & M32
    for i in range(1, N):
        E, O = (o, e) if i & 1 else (e, o)
        shift_mad_from(0, E, O, a[1:], b[i])
        O[N - 1] = (O[N - 1] + cmad_from(0, E, a, b[i])) & M32
        m = (E[0] * np0) & M32
        cmad_from(0, O, pl[1:], m)
        O[N - 1] = (O[N - 1] + cmad_from(0, E, pl, m)) & M32
    merge(e, o)
    r = value(e)
    return r - p if r >= p else r


def test_product_model_matches_definition_for_reduced_operands():
    rnd = random.Random(4)
    for p in (Q, RR):
        rinv = pow(1 << 256, -1, p)
        edge = [0, 1, 2, p - 1, p - 2, (1 << 253), 0xFFFFFFFF, 1 << 32, (1 << 224) - 1]
        for a, b in [(x, y) for x in edge for y in edge] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(400)]:
            assert mul_model(a, b, p) == a * b * rinv % p, (hex(a), hex(b))


def test_product_model_with_a_raw_operand_it_must_be_the_second_one():
    """to_mont(a) is how the witness enters the H pipeline (prover.cu::k_fr_to_mont), and the reference feeds
    frm_toMontgomery RAW 256-bit signals (src/bn128.js:134; a witness value need not be < r).  In the device product the
    running value of the row loop stays below A + p and must fit the nine limbs the two staggered accumulators span:
      mul(R^2, a)  -- raw value second: canonical a R mod p for EVERY 256-bit a (what fp.cuh::to_mont does);
      mul(a, R^2)  -- raw value first: right below 2^256 - p, loses a carry now and then above (what to_mont did until the
                      model found it: wrong for about 1 % of the witness values above 0.85 * 2^256 -- none occur in real
                      witnesses, all are legal inputs of the reference).
    The emulator cannot show this (it runs the portable host path); tests/test_ptx_kernels.py executes the compiled kernel."""
    rnd = random.Random(5)
    for p in (Q, RR):
        r2 = (1 << 512) % p
        rinv = pow(1 << 256, -1, p)
        top = (1 << 256) - 1
        cases = [top, top - 1, p, p + 1, 2 * p, 2 * p + 1, 5 * p + 3, top - p, top - p + 1, (1 << 255), (1 << 255) + 1, top ^ 0xFFFFFFFF, top ^ (0xFFFFFFFF << 224)]
        cases += [rnd.randrange(4 * p, 1 << 256) for _ in range(600)] + [rnd.randrange(1 << 256) for _ in range(200)]
        for a in cases:
            assert mul_model(r2, a, p) == (a << 256) % p, hex(a)
        for _ in range(200):                      # any reduced first operand, any raw second operand
            a, b = rnd.choice([p - 1, p - 2, rnd.randrange(p)]), rnd.choice([top, rnd.randrange(1 << 256)])
            assert mul_model(a, b, p) == a * b * rinv % p, (hex(a), hex(b))
        # the other order: exact while A + p < 2^256 ...
        for a in [p, 2 * p + 1, top - p - 1] + [rnd.randrange(p, (1 << 256) - p) for _ in range(200)]:
            assert mul_model(a, r2, p) == (a << 256) % p, hex(a)
        # ... and not above (the limit is real, not an artefact of a cautious bound)
        assert any(mul_model(a, r2, p) != (a << 256) % p for a in [top - 1, top ^ 0xFFFFFFFF] + [rnd.randrange(4 * p, 1 << 256) for _ in range(400)])
