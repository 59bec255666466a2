#!/usr/bin/env python3
"""Generates wasmsnark_b200/csrc/fp_asm.cuh: PTX carry-chain primitives for N-limb
(32-bit limbs) Montgomery arithmetic, N = 8 (BN128 / BLS12-381 Fr, 256 bit), N = 12 (BLS12-381 Fq, 384 bit) and N = 24 (MNT6753, 768 bit).

Every chain is ONE asm statement, so the carry flag never has to survive between
statements.  ptxas fuses each mad.lo.cc/madc.hi.cc pair into a single IMAD.WIDE.U32(.X):
the hardware multiplier is 32x32+64 -> 64, one SASS instruction per limb product.

The multiplication uses two accumulators, `e` aligned on limb 0 and `o` aligned on limb 1,
so that every product a[j]*b_i (which spans limbs j, j+1) lands on an aligned limb pair of
one of them and each row is a single uninterrupted carry chain.
"""
import sys


def gen(N):
    H = N // 2
    L = []
    w = L.append
    sfx = str(N)

    # ---- mul_n: acc[2k],acc[2k+1] = a[2k]*bi (a pre-offset by caller) ----
    w("__device__ __forceinline__ void mul_n%s(uint32_t* acc, const uint32_t* a, uint32_t bi) {" % sfx)
    w("#pragma unroll")
    w("    for (int k = 0; k < %d; k++) {" % H)
    w('        asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=&r"(acc[2 * k]), "=&r"(acc[2 * k + 1]) : "r"(a[2 * k]), "r"(bi));')
    w("    }")
    w("}")
    w("")

    # ---- cmad_n: acc[0..N) += sum_k a[2k]*bi << 64k ; returns carry out ----
    w("__device__ __forceinline__ uint32_t cmad_n%s(uint32_t* acc, const uint32_t* a, uint32_t bi) {" % sfx)
    w("    uint32_t c;")
    ins = []
    # operands: 0..N-1 acc, N = c, N+1..N+H a, N+H+1 bi
    for k in range(H):
        lo = "mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32"
        ins.append("%s %%%d, %%%d, %%%d, %%%d;" % (lo, 2 * k, N + 1 + k, N + 1 + H, 2 * k))
        ins.append("madc.hi.cc.u32 %%%d, %%%d, %%%d, %%%d;" % (2 * k + 1, N + 1 + k, N + 1 + H, 2 * k + 1))
    ins.append("addc.u32 %%%d, 0, 0;" % N)
    w('    asm("%s"' % " ".join(ins))
    w("        : %s, \"=&r\"(c)" % ", ".join('"+r"(acc[%d])' % i for i in range(N)))
    w("        : %s, \"r\"(bi));" % ", ".join('"r"(a[%d])' % (2 * k) for k in range(H)))
    w("    return c;")
    w("}")
    w("")

    # ---- shift_mad_n: E[0] += O[1]; O[j] = O[j+2] + (a[2k]*bi chain) ; O[N-2],O[N-1] = a*bi + carry ----
    w("__device__ __forceinline__ void shift_mad_n%s(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {" % sfx)
    ins = []
    # operands: 0..N-1 O (rw), N = E0 (rw), N+1..N+H a, N+H+1 bi
    ins.append("add.cc.u32 %%%d, %%%d, %%1;" % (N, N))
    for k in range(H):
        last = k == H - 1
        s_lo = "0" if last else "%%%d" % (2 * k + 2)
        s_hi = "0" if last else "%%%d" % (2 * k + 3)
        ins.append("madc.lo.cc.u32 %%%d, %%%d, %%%d, %s;" % (2 * k, N + 1 + k, N + 1 + H, s_lo))
        hi = "madc.hi.u32" if last else "madc.hi.cc.u32"
        ins.append("%s %%%d, %%%d, %%%d, %s;" % (hi, 2 * k + 1, N + 1 + k, N + 1 + H, s_hi))
    w('    asm("%s"' % " ".join(ins))
    w("        : %s, \"+r\"(E[0])" % ", ".join('"+r"(O[%d])' % i for i in range(N)))
    w("        : %s, \"r\"(bi));" % ", ".join('"r"(a[%d])' % (2 * k) for k in range(H)))
    w("}")
    w("")

    # ---- merge_n: r[k] = e[k] + o[k+1], k < N-1 ; r[N-1] = e[N-1] + carry  (in place on e) ----
    w("__device__ __forceinline__ void merge_n%s(uint32_t* e, const uint32_t* o) {" % sfx)
    ins = []
    for k in range(N - 1):
        op = "add.cc.u32" if k == 0 else "addc.cc.u32"
        ins.append("%s %%%d, %%%d, %%%d;" % (op, k, k, N + k))
    ins.append("addc.u32 %%%d, %%%d, 0;" % (N - 1, N - 1))
    w('    asm("%s"' % " ".join(ins))
    w("        : %s" % ", ".join('"+r"(e[%d])' % i for i in range(N)))
    w("        : %s);" % ", ".join('"r"(o[%d])' % (i + 1) for i in range(N - 1)))
    w("}")
    w("")

    # ---- add_n: r = a + b (no carry out: modulus has < 32N-1 bits) ----
    w("__device__ __forceinline__ void add_n%s(uint32_t* r, const uint32_t* a, const uint32_t* b) {" % sfx)
    ins = []
    for k in range(N):
        op = "add.cc.u32" if k == 0 else ("addc.u32" if k == N - 1 else "addc.cc.u32")
        ins.append("%s %%%d, %%%d, %%%d;" % (op, k, N + k, 2 * N + k))
    w('    asm("%s"' % " ".join(ins))
    w("        : %s" % ", ".join('"=&r"(r[%d])' % i for i in range(N)))
    w("        : %s, %s);" % (", ".join('"r"(a[%d])' % i for i in range(N)), ", ".join('"r"(b[%d])' % i for i in range(N))))
    w("}")
    w("")

    # ---- sub_n: r = a - b, returns borrow mask (0 or 0xffffffff) ----
    w("__device__ __forceinline__ uint32_t sub_n%s(uint32_t* r, const uint32_t* a, const uint32_t* b) {" % sfx)
    w("    uint32_t bw;")
    ins = []
    for k in range(N):
        op = "sub.cc.u32" if k == 0 else "subc.cc.u32"
        ins.append("%s %%%d, %%%d, %%%d;" % (op, k, N + 1 + k, 2 * N + 1 + k))
    ins.append("subc.u32 %%%d, 0, 0;" % N)
    w('    asm("%s"' % " ".join(ins))
    w("        : %s, \"=&r\"(bw)" % ", ".join('"=&r"(r[%d])' % i for i in range(N)))
    w("        : %s, %s);" % (", ".join('"r"(a[%d])' % i for i in range(N)), ", ".join('"r"(b[%d])' % i for i in range(N))))
    w("    return bw;")
    w("}")
    w("")

    # ---- addmask_n: r += (b & mask) ----
    w("__device__ __forceinline__ void addmask_n%s(uint32_t* r, const uint32_t* b, uint32_t mask) {" % sfx)
    w("    uint32_t t[%d];" % N)
    w("#pragma unroll")
    w("    for (int k = 0; k < %d; k++) t[k] = b[k] & mask;" % N)
    ins = []
    for k in range(N):
        op = "add.cc.u32" if k == 0 else ("addc.u32" if k == N - 1 else "addc.cc.u32")
        ins.append("%s %%%d, %%%d, %%%d;" % (op, k, k, N + k))
    w('    asm("%s"' % " ".join(ins))
    w("        : %s" % ", ".join('"+r"(r[%d])' % i for i in range(N)))
    w("        : %s);" % ", ".join('"r"(t[%d])' % i for i in range(N)))
    w("}")
    w("")
    if N == 8:
        gen_from_variants(N, w)
    return "\n".join(L)


def shift_mad_from_ins(N, K):
    """shift_mad_n with the products of pairs k < K known to be zero (the dedicated squaring skips a_j * a_i, j < i):
    those positions only move O down by two limbs and carry.  Operands: 0..N-1 O, N = E0, N+1.. = a[2k] (k >= K), last = bi."""
    H = N // 2
    ins = ["add.cc.u32 %%%d, %%%d, %%1;" % (N, N)]
    bi = N + 1 + (H - K)
    for k in range(H):
        last = k == H - 1
        s_lo = "0" if last else "%%%d" % (2 * k + 2)
        s_hi = "0" if last else "%%%d" % (2 * k + 3)
        if k < K:
            ins.append("addc.cc.u32 %%%d, %s, 0;" % (2 * k, s_lo))
            ins.append("addc.cc.u32 %%%d, %s, 0;" % (2 * k + 1, s_hi))
        else:
            a = N + 1 + (k - K)
            ins.append("madc.lo.cc.u32 %%%d, %%%d, %%%d, %s;" % (2 * k, a, bi, s_lo))
            ins.append("%s %%%d, %%%d, %%%d, %s;" % ("madc.hi.u32" if last else "madc.hi.cc.u32", 2 * k + 1, a, bi, s_hi))
    return ins


def cmad_from_ins(N, K):
    """cmad_n starting at pair K (lower limbs untouched).  Operands: 0..N-1 acc, N = carry out, N+1.. = a[2k] (k >= K), last = bi."""
    H = N // 2
    bi = N + 1 + (H - K)
    ins = []
    for k in range(K, H):
        a = N + 1 + (k - K)
        ins.append("%s %%%d, %%%d, %%%d, %%%d;" % ("mad.lo.cc.u32" if k == K else "madc.lo.cc.u32", 2 * k, a, bi, 2 * k))
        ins.append("madc.hi.cc.u32 %%%d, %%%d, %%%d, %%%d;" % (2 * k + 1, a, bi, 2 * k + 1))
    ins.append("addc.u32 %%%d, 0, 0;" % N)
    return ins


def gen_from_variants(N, w):
    H = N // 2
    sfx = str(N)
    w("// ---- variants for the dedicated squaring: products of pairs k < K are known to be zero ----")
    w("template <int K> __device__ __forceinline__ void shift_mad_from_n%s(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi);" % sfx)
    w("template <int K> __device__ __forceinline__ uint32_t cmad_from_n%s(uint32_t* acc, const uint32_t* a, uint32_t bi);" % sfx)
    for K in range(H):
        w("template <> __device__ __forceinline__ void shift_mad_from_n%s<%d>(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {" % (sfx, K))
        w('    asm("%s"' % " ".join(shift_mad_from_ins(N, K)))
        w("        : %s, \"+r\"(E[0])" % ", ".join('"+r"(O[%d])' % i for i in range(N)))
        w("        : %s, \"r\"(bi));" % ", ".join('"r"(a[%d])' % (2 * k) for k in range(K, H)))
        w("}")
        w("template <> __device__ __forceinline__ uint32_t cmad_from_n%s<%d>(uint32_t* acc, const uint32_t* a, uint32_t bi) {" % (sfx, K))
        w("    uint32_t c;")
        w('    asm("%s"' % " ".join(cmad_from_ins(N, K)))
        w("        : %s, \"=&r\"(c)" % ", ".join('"+r"(acc[%d])' % i for i in range(N)))
        w("        : %s, \"r\"(bi));" % ", ".join('"r"(a[%d])' % (2 * k) for k in range(K, H)))
        w("    return c;")
        w("}")
    w("")


def main(path):
    s = ("// GENERATED by tools/gen_fp_asm.py -- do not edit.\n"
         "// PTX carry-chain primitives (device only) for N-limb Montgomery arithmetic.\n"
         "#pragma once\n#include <cstdint>\n\nnamespace ws {\n\n#ifdef __CUDA_ARCH__\n\n")
    for N in (8, 12, 24):
        s += "// ======================= N = %d =======================\n" % N
        s += gen(N) + "\n"
    s += "#endif  // __CUDA_ARCH__\n\n}  // namespace ws\n"
    open(path, "w").write(s)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "wasmsnark_b200/csrc/fp_asm.cuh")
