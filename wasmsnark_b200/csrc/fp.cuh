// fp.cuh -- N x 32-bit-limb Montgomery prime field, values canonical in [0, p).
//
// Replaces the reference's generated WASM field module
//   f1m_add/sub/neg      /root/reference/src/build_f1m.js:67-113
//   f1m_mul              /root/reference/src/build_f1m.js:235-436   (R = 2^(32N), :37-39)
//   f1m_square           /root/reference/src/build_f1m.js:439-737
//   f1m_toMontgomery     /root/reference/src/build_f1m.js:749-756   (mul by R^2)
//   f1m_fromMontgomery   /root/reference/src/build_f1m.js:758-770   (mul by 1)
//   f1m_inverse          /root/reference/src/build_f1m.js:772-783
// Results are byte-identical to the reference's because every routine returns the
// canonical representative (the reference ends each op with a conditional subtract,
// build_f1m.js:426-435), whatever the internal schedule.
//
// Device path: whole number held in registers, each row of the product is one PTX
// mad.lo.cc / madc.hi.cc carry chain (fp_asm.cuh) that ptxas lowers to one IMAD.WIDE.U32
// per limb product.  Host path (g++ / nvcc host pass): portable CIOS on 64-bit limbs, used by
// the CPU unit tests of the group-law code and by the host-side Groth16 verifier (pairing.cuh).
#pragma once
#include <cstdint>
#include "hd.cuh"
#include "field_consts.cuh"
#include "fp_asm.cuh"

namespace ws {

template <class P>
struct Fp {
    static constexpr int N = P::N;
    uint32_t v[N];

    HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = 0;
        return r;
    }
    HD static Fp one() {   // Montgomery form of 1  (f1m_one)
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::one(i);
        return r;
    }
    HD static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::r2(i);
        return r;
    }
    HD static Fp modulus() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.v[i] = P::mod(i);
        return r;
    }
    HD bool is_zero() const {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < N; i++) t |= v[i];
        return t == 0;
    }
    HD bool operator==(const Fp& o) const {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < N; i++) t |= v[i] ^ o.v[i];
        return t == 0;
    }
    HD bool operator!=(const Fp& o) const { return !(*this == o); }
};

// ---------------------------------------------------------------- add / sub / neg
template <class P>
HD Fp<P> add(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> s, t;
#ifdef __CUDA_ARCH__
    const Fp<P> m = Fp<P>::modulus();
    uint32_t bw;
    if constexpr (N == 8) {
        add_n8(s.v, a.v, b.v);
        bw = sub_n8(t.v, s.v, m.v);
    } else if constexpr (N == 12) {
        add_n12(s.v, a.v, b.v);
        bw = sub_n12(t.v, s.v, m.v);
    } else {
        add_n24(s.v, a.v, b.v);
        bw = sub_n24(t.v, s.v, m.v);
    }
#pragma unroll
    for (int i = 0; i < N; i++) s.v[i] = bw ? s.v[i] : t.v[i];
#else
    // host: 64-bit limbs
    constexpr int M = N / 2;
    typedef unsigned __int128 u128;
    uint64_t sw[M], tw[M];
    u128 c = 0;
    for (int i = 0; i < M; i++) {
        c += (u128)(a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32)) + (b.v[2 * i] | ((uint64_t)b.v[2 * i + 1] << 32));
        sw[i] = (uint64_t)c;
        c >>= 64;
    }
    uint64_t bw = 0;
    for (int i = 0; i < M; i++) {
        u128 d = (u128)sw[i] - (P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32)) - bw;
        tw[i] = (uint64_t)d;
        bw = (uint64_t)(d >> 64) & 1;
    }
    for (int i = 0; i < M; i++) {
        uint64_t w = bw ? sw[i] : tw[i];
        s.v[2 * i] = (uint32_t)w;
        s.v[2 * i + 1] = (uint32_t)(w >> 32);
    }
    (void)t;
#endif
    return s;
}

template <class P>
HD Fp<P> sub(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> d;
#ifdef __CUDA_ARCH__
    const Fp<P> m = Fp<P>::modulus();
    if constexpr (N == 8) {
        uint32_t bw = sub_n8(d.v, a.v, b.v);
        addmask_n8(d.v, m.v, bw);
    } else if constexpr (N == 12) {
        uint32_t bw = sub_n12(d.v, a.v, b.v);
        addmask_n12(d.v, m.v, bw);
    } else {
        uint32_t bw = sub_n24(d.v, a.v, b.v);
        addmask_n24(d.v, m.v, bw);
    }
#else
    constexpr int M = N / 2;
    typedef unsigned __int128 u128;
    uint64_t dw[M], bw = 0;
    for (int i = 0; i < M; i++) {
        u128 t = (u128)(a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32)) - (b.v[2 * i] | ((uint64_t)b.v[2 * i + 1] << 32)) - bw;
        dw[i] = (uint64_t)t;
        bw = (uint64_t)(t >> 64) & 1;
    }
    if (bw) {
        u128 c = 0;
        for (int i = 0; i < M; i++) {
            c += (u128)dw[i] + (P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32));
            dw[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    for (int i = 0; i < M; i++) {
        d.v[2 * i] = (uint32_t)dw[i];
        d.v[2 * i + 1] = (uint32_t)(dw[i] >> 32);
    }
#endif
    return d;
}

template <class P>
HD Fp<P> neg(const Fp<P>& a) {
    // f1m_neg: 0 -> 0, else p - a   (build_f1m.js:103-113)
    if (a.is_zero()) return a;
    return sub(Fp<P>::zero(), a);
}

template <class P>
HD Fp<P> dbl(const Fp<P>& a) { return add(a, a); }

// ---------------------------------------------------------------- Montgomery product
template <class P>
HD Fp<P> mul_inline(const Fp<P>& A, const Fp<P>& B) {
    constexpr int N = P::N;
    Fp<P> res;
#ifdef __CUDA_ARCH__
    // Two accumulators: e aligned on limb 0, o aligned on limb 1 (see tools/gen_fp_asm.py).
    // After each row the running value is divided by 2^32, which swaps their roles.
    uint32_t e[N], o[N];
    const Fp<P> M = Fp<P>::modulus();
    const uint32_t* a = A.v;
    const uint32_t* b = B.v;
    const uint32_t* p = M.v;
#define WS_ROW0(NN)                                   \
    mul_n##NN(e, a, b[0]);                            \
    mul_n##NN(o, a + 1, b[0]);                        \
    {                                                 \
        uint32_t m = e[0] * P::NP0;                   \
        cmad_n##NN(o, p + 1, m);                      \
        o[N - 1] += cmad_n##NN(e, p, m);              \
    }
#define WS_ROW(NN, E, O, bi)                          \
    {                                                 \
        shift_mad_n##NN(E, O, a + 1, bi);             \
        O[N - 1] += cmad_n##NN(E, a, bi);             \
        uint32_t m = E[0] * P::NP0;                   \
        cmad_n##NN(O, p + 1, m);                      \
        O[N - 1] += cmad_n##NN(E, p, m);              \
    }
    if constexpr (N == 8) {
        WS_ROW0(8)
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            WS_ROW(8, o, e, b[i])
            if (i + 1 < N) WS_ROW(8, e, o, b[i + 1])
        }
        merge_n8(e, o);
        uint32_t bw = sub_n8(res.v, e, p);
#pragma unroll
        for (int i = 0; i < N; i++) res.v[i] = bw ? e[i] : res.v[i];
    } else if constexpr (N == 12) {
        WS_ROW0(12)
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            WS_ROW(12, o, e, b[i])
            if (i + 1 < N) WS_ROW(12, e, o, b[i + 1])
        }
        merge_n12(e, o);
        uint32_t bw = sub_n12(res.v, e, p);
#pragma unroll
        for (int i = 0; i < N; i++) res.v[i] = bw ? e[i] : res.v[i];
    } else {
        WS_ROW0(24)
#pragma unroll
        for (int i = 1; i < N; i += 2) {
            WS_ROW(24, o, e, b[i])
            if (i + 1 < N) WS_ROW(24, e, o, b[i + 1])
        }
        merge_n24(e, o);
        uint32_t bw = sub_n24(res.v, e, p);
#pragma unroll
        for (int i = 0; i < N; i++) res.v[i] = bw ? e[i] : res.v[i];
    }
#undef WS_ROW0
#undef WS_ROW
#else
    // portable CIOS on 64-bit limbs (host: unit tests of the group-law code, the Groth16 verifier)
    constexpr int M = N / 2;
    static_assert(N % 2 == 0, "even limb count");
    typedef unsigned __int128 u128;
    uint64_t a[M], b[M], p[M], t[M + 2];
    for (int i = 0; i < M; i++) {
        a[i] = A.v[2 * i] | ((uint64_t)A.v[2 * i + 1] << 32);
        b[i] = B.v[2 * i] | ((uint64_t)B.v[2 * i + 1] << 32);
        p[i] = P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
    }
    uint64_t pinv = (uint64_t)(0u - P::NP0);            // p^-1 mod 2^32, lifted to 2^64 by one Newton step
    pinv *= 2 - p[0] * pinv;
    const uint64_t np = 0 - pinv;
    for (int i = 0; i < M + 2; i++) t[i] = 0;
    for (int i = 0; i < M; i++) {
        u128 c = 0;
        for (int j = 0; j < M; j++) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[M];
        t[M] = (uint64_t)c;
        t[M + 1] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * np;
        c = ((u128)m * p[0] + t[0]) >> 64;
        for (int j = 1; j < M; j++) {
            c += (u128)m * p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[M];
        t[M - 1] = (uint64_t)c;
        t[M] = t[M + 1] + (uint64_t)(c >> 64);
    }
    uint64_t d[M], bw = 0;
    for (int i = 0; i < M; i++) {
        u128 x = (u128)t[i] - p[i] - bw;
        d[i] = (uint64_t)x;
        bw = (uint64_t)(x >> 64) & 1;
    }
    bool ge = t[M] != 0 || bw == 0;
    for (int i = 0; i < M; i++) {
        uint64_t w = ge ? d[i] : t[i];
        res.v[2 * i] = (uint32_t)w;
        res.v[2 * i + 1] = (uint32_t)(w >> 32);
    }
#endif
    return res;
}

// Big products go out of line -- one copy per kernel instead of one per call site keeps ptxas time and code size
// sane: always for 24 limbs (~1300 instructions); for 12 limbs (~330) only under the quadratic extension (mul_ool,
// used by fp2.cuh: a G2 point addition would inline 30 of them).  The 8-limb product stays inline (the hot path).
template <class P>
HD_NOINLINE void mul_to(Fp<P>* r, const Fp<P>* a, const Fp<P>* b) { *r = mul_inline(*a, *b); }

template <class P>
HD Fp<P> mul_ool(const Fp<P>& a, const Fp<P>& b) {
    if constexpr (P::N > 8) {
        Fp<P> r;
        mul_to<P>(&r, &a, &b);
        return r;
    } else {
        return mul_inline(a, b);
    }
}

template <class P>
HD Fp<P> mul(const Fp<P>& a, const Fp<P>& b) {
    if constexpr (P::N > 12) return mul_ool(a, b);
    else return mul_inline(a, b);
}

// Dedicated squaring for the 8-limb fields (device): a_i a_j, i != j, is computed once and doubled, so the product
// part needs 36 limb products instead of 64 (the 72 of the reduction stay): 108 instead of 137 IMAD-class instructions.
// Row i of the interleaved scheme multiplies a_i by the vector (0, .., 0, a_i, 2 a_{>i}) -- the limbs of the doubled
// number d = 2a, with the bit that a_i shifts into d_{i+1} cleared -- through the chain variants that skip the zero
// products (fp_asm.cuh: shift_mad_from / cmad_from).  Needs two spare bits in the top limb (p < 2^254: the doubled
// early rows stay below 2^288, the nine limbs of the two accumulators; a 255-bit modulus such as BLS12-381 Fr overflows
// them, so it keeps sqr = mul).
// tests/test_fp_asm_model.py runs this composition on the generated instruction lists in a Python model.
#ifdef __CUDA_ARCH__
template <class P>
__device__ __forceinline__ Fp<P> sqr_inline8(const Fp<P>& A) {
    constexpr int N = 8;
    static_assert(P::N == N && P::BITS <= 254, "8 limbs with two spare top bits");
    Fp<P> res;
    uint32_t e[N], o[N], d[N];
    const Fp<P> M = Fp<P>::modulus();
    const uint32_t* a = A.v;
    const uint32_t* p = M.v;
    d[0] = 0;
#pragma unroll
    for (int j = 1; j < N; j++) d[j] = __funnelshift_l(a[j - 1], a[j], 1);
#define WS_SQR_X(i)                                                                                   \
    uint32_t x[N + 1];                                                                                \
    _Pragma("unroll") for (int j = 0; j < N; j++)                                                     \
        x[j] = j < (i) ? 0u : (j == (i) ? a[(i)] : (j == (i) + 1 ? (d[j] & ~1u) : d[j]));             \
    x[N] = 0;
#define WS_SQR_REDUCE(E, O)                                                                           \
    {                                                                                                 \
        uint32_t m = E[0] * P::NP0;                                                                   \
        cmad_n8(O, p + 1, m);                                                                         \
        O[N - 1] += cmad_n8(E, p, m);                                                                 \
    }
#define WS_SQR_ROW(i, E, O)                                                                           \
    {                                                                                                 \
        WS_SQR_X(i)                                                                                   \
        shift_mad_from_n8<((i) >> 1)>(E, O, x + 1, a[(i)]);                                           \
        if constexpr ((((i) + 1) >> 1) < N / 2) O[N - 1] += cmad_from_n8<(((i) + 1) >> 1)>(E, x, a[(i)]); \
        WS_SQR_REDUCE(E, O)                                                                           \
    }
    {
        WS_SQR_X(0)
        mul_n8(e, x, a[0]);
        mul_n8(o, x + 1, a[0]);
        WS_SQR_REDUCE(e, o)
    }
    WS_SQR_ROW(1, o, e)
    WS_SQR_ROW(2, e, o)
    WS_SQR_ROW(3, o, e)
    WS_SQR_ROW(4, e, o)
    WS_SQR_ROW(5, o, e)
    WS_SQR_ROW(6, e, o)
    WS_SQR_ROW(7, o, e)
#undef WS_SQR_X
#undef WS_SQR_REDUCE
#undef WS_SQR_ROW
    merge_n8(e, o);
    uint32_t bw = sub_n8(res.v, e, p);
#pragma unroll
    for (int i = 0; i < N; i++) res.v[i] = bw ? e[i] : res.v[i];
    return res;
}
#endif

template <class P>
HD Fp<P> sqr(const Fp<P>& a) {
#if defined(__CUDA_ARCH__) && !defined(WS_NO_DEDICATED_SQR)
    if constexpr (P::N == 8 && P::BITS <= 254) return sqr_inline8(a);
    else return mul(a, a);
#else
    return mul(a, a);
#endif
}

// f1m_toMontgomery (build_f1m.js:118-125) takes RAW N-limb numbers -- a witness value need not be below the modulus
// (src/bn128.js:134 hands frm_toMontgomeryN the signals as they come) -- and returns the canonical a R mod p.  The raw
// value must be the SECOND operand of the device product: the row loop's running value stays below A + p, which has to fit
// the N + 1 limbs the two staggered accumulators span (A + p < 2^(32 N)); B only ever enters limb by limb.  With the raw
// value first, a >= 2^256 - p loses a carry now and then (tests/test_fp_asm_model.py shows both orders on the generated
// PTX, tests/test_ptx_kernels.py executes the compiled kernels).
template <class P>
HD Fp<P> to_mont(const Fp<P>& a) { return mul(Fp<P>::r2(), a); }

template <class P>
HD Fp<P> from_mont(const Fp<P>& a) {
    Fp<P> one = Fp<P>::zero();
    one.v[0] = 1;
    return mul(a, one);
}

// a^(p-2): the reference inverts with int_inverseMod (build_f1m.js:772-783); the inverse of
// a non-zero residue is unique, so Fermat gives the same canonical bytes.  inverse(0) = 0.
template <class P>
HD Fp<P> inv(const Fp<P>& a) {
    constexpr int N = P::N;
    Fp<P> r = Fp<P>::one();
    bool started = false;
    for (int i = N - 1; i >= 0; i--) {
        uint32_t w = P::pm2(i);
        for (int b = 31; b >= 0; b--) {
            if (started) r = sqr(r);
            if ((w >> b) & 1) {
                r = started ? mul(r, a) : a;
                started = true;
            }
        }
    }
    return r;
}

template <class P>
HD Fp<P> select(bool c, const Fp<P>& a, const Fp<P>& b) {   // c ? a : b
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < P::N; i++) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
}

using Fq = Fp<BN128_Fq>;
using Fr = Fp<BN128_Fr>;

}  // namespace ws
