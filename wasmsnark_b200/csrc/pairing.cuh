// pairing.cuh -- BN128 optimal-ate pairing on the HOST (the Groth16 verifier is four pairings: not data-parallel,
// SURVEY.md §8f rank 3 keeps it on the CPU).  Built on the same Fp / Fp2 classes as the device code (their portable
// host paths).
//
// Replaces, in the reference's bn128 image:
//   bn128_prepareG1 / prepareG2 / millerLoop     /root/reference/src/bn128/build_bn128.js:265-836
//   bn128__frobeniusMap{1,2,3}                   /root/reference/src/bn128/build_bn128.js:838-941
//   bn128_finalExponentiation                    /root/reference/src/bn128/build_bn128.js:943-1262
//   bn128_pairing / bn128_pairingEq{1..5}        /root/reference/src/bn128/build_bn128.js:1287-1346
// The reference precomputes projective line coefficients for the G2 point; here the Miller loop walks the twist
// point in affine coordinates (one Fq2 inversion per step is cheap on a CPU and keeps the lines three-term).  The
// Miller values therefore differ from the reference's by factors in proper subfields of Fq12, which the final
// exponentiation removes: the pairing VALUE (bn128_pairing's output) and every verdict are identical.
// Final exponent, as the reference: (q^6-1)(q^2+1) then the hard part raised to
//   q^3 (12z^3+6z^2+4z-1) + q^2 (12z^3+6z^2+6z) + q (12z^3+6z^2+4z) + (12z^3+12z^2+6z+1)     (:1154-1243)
#pragma once
#include <cstring>
#include "ec.cuh"
#include "fp2.cuh"
#include "pairing_consts.cuh"

namespace ws {
namespace pairing {

using F2 = Fq2;

static inline Fq fq_c(const uint32_t v[8]) {
    Fq r;
    for (int i = 0; i < 8; i++) r.v[i] = v[i];
    return r;
}
static inline F2 f2_c(const uint32_t c[2][8]) { return F2{fq_c(c[0]), fq_c(c[1])}; }
static inline F2 conj(const F2& a) { return F2{a.c0, neg(a.c1)}; }
static inline F2 scale(const F2& a, const Fq& k) { return F2{mul(a.c0, k), mul(a.c1, k)}; }
// (9 + u)(a0 + a1 u) = 9 a0 - a1 + (9 a1 + a0) u; 9 x as one Montgomery product by the constant 9 (cheaper on the
// host than the doubling chain, whose every step ends in a conditional subtraction)
static inline F2 mul_xi(const F2& a) {
    static const Fq nine = fq_c(pairing_consts::XI[0]);
    return F2{sub(mul(a.c0, nine), a.c1), add(mul(a.c1, nine), a.c0)};
}

// Inverse of a Montgomery residue x = a R by the binary "almost inverse" (Kaliski): phase 1 yields x^-1 2^k mod p with
// 256 <= k <= 512 using only shifts and subtractions; the result a^-1 R = x^-1 R^2 = (x^-1 2^k) 2^(512-k) follows by
// modular doublings.  ~10x cheaper on a CPU than the a^(p-2) ladder the kernels use; inverse(0) = 0 like f1m_inverse.
struct U256 {
    uint64_t w[4];
};
static inline bool u_is_zero(const U256& a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) == 0; }
static inline bool u_gt(const U256& a, const U256& b) {
    for (int i = 3; i >= 0; i--)
        if (a.w[i] != b.w[i]) return a.w[i] > b.w[i];
    return false;
}
static inline uint64_t u_add(U256& a, const U256& b) {   // a += b, returns carry
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (unsigned __int128)a.w[i] + b.w[i];
        a.w[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static inline void u_sub(U256& a, const U256& b) {       // a -= b (a >= b)
    unsigned __int128 bw = 0;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 d = (unsigned __int128)a.w[i] - b.w[i] - bw;
        a.w[i] = (uint64_t)d;
        bw = (uint64_t)(d >> 64) & 1;
    }
}
static inline void u_shr1(U256& a) {
    for (int i = 0; i < 3; i++) a.w[i] = (a.w[i] >> 1) | (a.w[i + 1] << 63);
    a.w[3] >>= 1;
}
static inline uint64_t u_shl1(U256& a) {                 // returns the bit shifted out
    uint64_t out = a.w[3] >> 63;
    for (int i = 3; i > 0; i--) a.w[i] = (a.w[i] << 1) | (a.w[i - 1] >> 63);
    a.w[0] <<= 1;
    return out;
}
static inline Fq fq_inv_fast(const Fq& x) {
    U256 p, u, v, r{{0, 0, 0, 0}}, s{{1, 0, 0, 0}};
    for (int i = 0; i < 4; i++) {
        p.w[i] = BN128_Fq::mod(2 * i) | ((uint64_t)BN128_Fq::mod(2 * i + 1) << 32);
        v.w[i] = x.v[2 * i] | ((uint64_t)x.v[2 * i + 1] << 32);
    }
    if (u_is_zero(v)) return x;
    u = p;
    int k = 0;
    // invariants: x r = -u 2^k, x s = v 2^k (mod p); r, s < 2p fit 256 bits because p < 2^254
    while (!u_is_zero(v)) {
        if (!(u.w[0] & 1)) {
            u_shr1(u);
            u_shl1(s);
        } else if (!(v.w[0] & 1)) {
            u_shr1(v);
            u_shl1(r);
        } else if (u_gt(u, v)) {
            u_sub(u, v);
            u_shr1(u);
            u_add(r, s);
            u_shl1(s);
        } else {
            u_sub(v, u);
            u_shr1(v);
            u_add(s, r);
            u_shl1(r);
        }
        k++;
    }
    if (!u_gt(p, r)) u_sub(r, p);           // r in [0, 2p) -> [0, p)
    U256 t = p;
    u_sub(t, r);                            // x^-1 2^k = p - r
    for (; k < 512; k++) {                  // times 2^(512 - k)
        u_shl1(t);                          // t < p < 2^254: no bit is lost
        if (!u_gt(p, t)) u_sub(t, p);
    }
    Fq out;
    for (int i = 0; i < 4; i++) {
        out.v[2 * i] = (uint32_t)t.w[i];
        out.v[2 * i + 1] = (uint32_t)(t.w[i] >> 32);
    }
    return out;
}
static inline F2 f2_inv_fast(const F2& x) {              // 1 / (c0^2 + c1^2) * (c0, -c1)   (src/build_f2m.js:275-333)
    Fq t = fq_inv_fast(add(sqr(x.c0), sqr(x.c1)));
    return F2{mul(x.c0, t), neg(mul(x.c1, t))};
}

// ---------------------------------------------------------------- Fq6 = Fq2[v]/(v^3 - xi)
struct F6 {
    F2 a0, a1, a2;
    static F6 zero() { return F6{F2::zero(), F2::zero(), F2::zero()}; }
    static F6 one() { return F6{F2::one(), F2::zero(), F2::zero()}; }
    bool operator==(const F6& o) const { return a0 == o.a0 && a1 == o.a1 && a2 == o.a2; }
};
static inline F6 add(const F6& x, const F6& y) { return F6{add(x.a0, y.a0), add(x.a1, y.a1), add(x.a2, y.a2)}; }
static inline F6 sub(const F6& x, const F6& y) { return F6{sub(x.a0, y.a0), sub(x.a1, y.a1), sub(x.a2, y.a2)}; }
static inline F6 neg(const F6& x) { return F6{neg(x.a0), neg(x.a1), neg(x.a2)}; }
static inline F6 mul_v(const F6& x) { return F6{mul_xi(x.a2), x.a0, x.a1}; }
static inline F6 mul(const F6& x, const F6& y) {
    F2 t0 = mul(x.a0, y.a0), t1 = mul(x.a1, y.a1), t2 = mul(x.a2, y.a2);
    F2 m12 = sub(sub(mul(add(x.a1, x.a2), add(y.a1, y.a2)), t1), t2);
    F2 m01 = sub(sub(mul(add(x.a0, x.a1), add(y.a0, y.a1)), t0), t1);
    F2 m02 = sub(sub(mul(add(x.a0, x.a2), add(y.a0, y.a2)), t0), t2);
    return F6{add(t0, mul_xi(m12)), add(m01, mul_xi(t2)), add(m02, t1)};
}
static inline F6 inv(const F6& x) {
    F2 A = sub(sqr(x.a0), mul_xi(mul(x.a1, x.a2)));
    F2 B = sub(mul_xi(sqr(x.a2)), mul(x.a0, x.a1));
    F2 C = sub(sqr(x.a1), mul(x.a0, x.a2));
    F2 t = add(mul(x.a0, A), mul_xi(add(mul(x.a2, B), mul(x.a1, C))));
    F2 ti = f2_inv_fast(t);
    return F6{mul(A, ti), mul(B, ti), mul(C, ti)};
}

// ---------------------------------------------------------------- Fq12 = Fq6[w]/(w^2 - v)
struct F12 {
    F6 c0, c1;
    static F12 one() { return F12{F6::one(), F6::zero()}; }
    bool operator==(const F12& o) const { return c0 == o.c0 && c1 == o.c1; }
};
static inline F12 mul(const F12& x, const F12& y) {
    F6 t0 = mul(x.c0, y.c0), t1 = mul(x.c1, y.c1);
    F6 m = sub(sub(mul(add(x.c0, x.c1), add(y.c0, y.c1)), t0), t1);
    return F12{add(t0, mul_v(t1)), m};
}
static inline F12 sqr(const F12& x) {
    // (c0 + c1 w)^2 = c0^2 + v c1^2 + 2 c0 c1 w, with two Fq6 products
    F6 p = mul(x.c0, x.c1);
    F6 s = mul(add(x.c0, x.c1), add(x.c0, mul_v(x.c1)));
    return F12{sub(sub(s, p), mul_v(p)), add(p, p)};
}
static inline F12 conj(const F12& x) { return F12{x.c0, neg(x.c1)}; }
static inline F12 inv(const F12& x) {
    F6 t = sub(mul(x.c0, x.c0), mul_v(mul(x.c1, x.c1)));
    F6 ti = inv(t);
    return F12{mul(x.c0, ti), neg(mul(x.c1, ti))};
}
// coefficient of w^i, i = 0..5  (w^2 = v):  c0.a0, c1.a0, c0.a1, c1.a1, c0.a2, c1.a2
static inline F2& coeff(F12& x, int i) {
    F6& h = (i & 1) ? x.c1 : x.c0;
    return (i >> 1) == 0 ? h.a0 : (i >> 1) == 1 ? h.a1 : h.a2;
}
// x^(q^k), k = 1..3: (a_i w^i)^(q^k) = frob^k(a_i) * xi^(i (q^k - 1)/6) * w^i
static inline F12 frobenius(const F12& x, int k) {
    const uint32_t(*g)[2][8] = k == 1 ? pairing_consts::GAMMA1 : k == 2 ? pairing_consts::GAMMA2 : pairing_consts::GAMMA3;
    F12 r = x;
    for (int i = 0; i < 6; i++) {
        F2 a = coeff(r, i);
        if (k & 1) a = conj(a);
        coeff(r, i) = mul(a, f2_c(g[i]));
    }
    return r;
}
// x^2 for x in the cyclotomic subgroup (everything after the easy part of the final exponentiation): Granger-Scott
// squaring -- three Fq4 squarings (z0 + z1 s, z2 + z3 s, z4 + z5 s with s^2 = xi), 6 Fq2 products instead of 12.
// Replaces bn128__cyclotomicSquare, /root/reference/src/bn128/build_bn128.js:980-1090.
static inline void fq4_sqr(const F2& a, const F2& b, F2* t0, F2* t1) {   // (a + b s)^2 = (a^2 + xi b^2) + 2ab s
    F2 ab = mul(a, b);
    *t0 = sub(sub(mul(add(a, b), add(a, mul_xi(b))), ab), mul_xi(ab));
    *t1 = dbl(ab);
}
static inline F12 cyclotomic_sqr(const F12& x) {
    F2 z0 = x.c0.a0, z4 = x.c0.a1, z3 = x.c0.a2, z2 = x.c1.a0, z1 = x.c1.a1, z5 = x.c1.a2;
    F2 t0, t1, t2, t3, t4, t5;
    fq4_sqr(z0, z1, &t0, &t1);
    fq4_sqr(z2, z3, &t2, &t3);
    fq4_sqr(z4, z5, &t4, &t5);
    F2 xt5 = mul_xi(t5);
    z0 = add(dbl(sub(t0, z0)), t0);       // 3 t0 - 2 z0
    z1 = add(dbl(add(t1, z1)), t1);       // 3 t1 + 2 z1
    z2 = add(dbl(add(xt5, z2)), xt5);     // 3 xi t5 + 2 z2
    z3 = add(dbl(sub(t4, z3)), t4);       // 3 t4 - 2 z3
    z4 = add(dbl(sub(t2, z4)), t2);       // 3 t2 - 2 z4
    z5 = add(dbl(add(t3, z5)), t3);       // 3 t3 + 2 z5
    return F12{F6{z0, z4, z3}, F6{z2, z1, z5}};
}
static inline F12 pow_z(const F12& x) {   // x^z, z = 4965661367192848881 (63 bits); x in the cyclotomic subgroup
    F12 r = x;
    unsigned long long z = pairing_consts::Z;
    int top = 63;
    while (!((z >> top) & 1)) top--;
    for (int b = top - 1; b >= 0; b--) {
        r = cyclotomic_sqr(r);
        if ((z >> b) & 1) r = mul(r, x);
    }
    return r;
}

// ---------------------------------------------------------------- points
struct G1A {
    Fq x, y;
    bool inf;
};
struct G2A {
    F2 x, y;
    bool inf;
};
static inline bool on_curve(const G1A& p) {   // y^2 = x^3 + 3
    if (p.inf) return true;
    Fq three = add(add(Fq::one(), Fq::one()), Fq::one());
    return sqr(p.y) == add(mul(sqr(p.x), p.x), three);
}
static inline bool on_curve(const G2A& p) {   // y^2 = x^3 + 3/xi
    if (p.inf) return true;
    return sqr(p.y) == add(mul(sqr(p.x), p.x), f2_c(pairing_consts::TWIST_B));
}

// (x0, x1, x2) * (A, B, 0) in Fq6: 5 Fq2 products
static inline F6 mul_sparse_ab(const F6& x, const F2& A, const F2& B) {
    F2 x0A = mul(x.a0, A), x1B = mul(x.a1, B);
    F2 c1 = sub(sub(mul(add(x.a0, x.a1), add(A, B)), x0A), x1B);      // x0 B + x1 A
    return F6{add(x0A, mul_xi(mul(x.a2, B))), c1, add(x1B, mul(x.a2, A))};
}
// f times the line through the (untwisted) point T with twist-slope lam, evaluated at P:
//   l = yP - lam xP w + (lam xT - yT) w^3 = (yP, 0, 0) + (A, B, 0) w,  A = -lam xP, B = lam xT - yT
// (the reference's bn128__mulBy024, /root/reference/src/bn128/build_bn128.js:613-754, plays this role): Karatsuba over
// w with both partial products sparse -- 6 Fq + 10 Fq2 products instead of 18 Fq2.
static inline F12 mul_by_line(const F12& f, const G2A& T, const F2& lam, const G1A& P) {
    F2 A = neg(scale(lam, P.x)), B = sub(mul(lam, T.x), T.y);
    F6 t0{scale(f.c0.a0, P.y), scale(f.c0.a1, P.y), scale(f.c0.a2, P.y)};
    F6 t1 = mul_sparse_ab(f.c1, A, B);
    F6 t2 = mul_sparse_ab(add(f.c0, f.c1), F2{add(A.c0, P.y), A.c1}, B);
    return F12{add(t0, mul_v(t1)), sub(sub(t2, t0), t1)};
}

struct MillerState {
    G1A P;
    G2A Q, T;
};
// d[i] <- 1/d[i] for n <= 8 values with one field inversion.  Returns false (d untouched) if any d[i] is zero: a zero
// would otherwise zero EVERY inverse of the batch, i.e. one degenerate pair would corrupt the slopes of the others.
static inline bool batch_inv(F2* d, int n) {
    F2 pre[8];
    F2 acc = F2::one();
    for (int i = 0; i < n; i++)
        if (d[i].is_zero()) return false;
    for (int i = 0; i < n; i++) {
        pre[i] = acc;
        acc = mul(acc, d[i]);
    }
    acc = f2_inv_fast(acc);
    for (int i = n - 1; i >= 0; i--) {
        F2 t = mul(acc, pre[i]);
        acc = mul(acc, d[i]);
        d[i] = t;
    }
    return true;
}
// T <- T + R along the chord/tangent of slope lam, multiplying the line into f
static inline void step_apply(MillerState& s, const F2& lam, const F2& rx, F12& f) {
    f = mul_by_line(f, s.T, lam, s.P);
    F2 x3 = sub(sub(sqr(lam), s.T.x), rx);
    s.T.y = sub(mul(lam, sub(s.T.x, x3)), s.T.y);
    s.T.x = x3;
}
// The affine walk has no formula for a vertical line: T = +-R in an addition step or y_T = 0 in a doubling step.  Neither
// can happen for Q in the order-r subgroup (the loop count stays below r); it can for a twist point outside it (the
// BN254 twist cofactor has small factors) -- the steps then report failure and the caller rejects the input.
static inline bool steps_double(MillerState** st, int n, F12& f) {
    F2 den[8];
    for (int i = 0; i < n; i++) den[i] = dbl(st[i]->T.y);
    if (!batch_inv(den, n)) return false;
    for (int i = 0; i < n; i++) {
        F2 x2 = sqr(st[i]->T.x);
        step_apply(*st[i], mul(add(dbl(x2), x2), den[i]), st[i]->T.x, f);
    }
    return true;
}
static inline bool steps_add(MillerState** st, const G2A* R, int n, F12& f) {
    F2 den[8];
    for (int i = 0; i < n; i++) den[i] = sub(R[i].x, st[i]->T.x);
    if (!batch_inv(den, n)) return false;
    for (int i = 0; i < n; i++) step_apply(*st[i], mul(sub(R[i].y, st[i]->T.y), den[i]), R[i].x, f);
    return true;
}

// prod_i f_{6z+2,Q_i}(P_i) * (the two Frobenius lines): one shared squaring and one shared inversion per step
// *ok = false: a degenerate step was met (some Q_i is not in the order-r subgroup); the returned value is meaningless.
static inline F12 miller_product(const G1A* P, const G2A* Q, int n, bool* ok) {
    *ok = true;
    MillerState all[8];
    MillerState* st[8];
    G2A q0[8], q1[8], q2[8];
    int m = 0;
    for (int i = 0; i < n && m < 8; i++) {
        if (P[i].inf || Q[i].inf) continue;                      // e(O, Q) = e(P, O) = 1
        all[m] = MillerState{P[i], Q[i], Q[i]};
        st[m] = &all[m];
        q0[m] = Q[i];
        q1[m] = G2A{mul(conj(Q[i].x), f2_c(pairing_consts::GAMMA1[2])), mul(conj(Q[i].y), f2_c(pairing_consts::GAMMA1[3])), false};   // pi(Q)
        q2[m] = G2A{mul(Q[i].x, f2_c(pairing_consts::GAMMA2[2])), neg(mul(Q[i].y, f2_c(pairing_consts::GAMMA2[3]))), false};          // -pi^2(Q)
        m++;
    }
    F12 f = F12::one();
    if (!m) return f;
    for (int b = 1; b < pairing_consts::ATE_LOOP_BITS; b++) {
        f = sqr(f);
        if (!steps_double(st, m, f) || (pairing_consts::ATE_LOOP[b] && !steps_add(st, q0, m, f))) {
            *ok = false;
            return f;
        }
    }
    if (!steps_add(st, q1, m, f) || !steps_add(st, q2, m, f)) *ok = false;
    return f;
}

static inline F12 final_exponentiation(const F12& f) {
    // easy part: f^((q^6 - 1)(q^2 + 1)); afterwards the inverse is the conjugate
    F12 e = mul(conj(f), inv(f));
    e = mul(frobenius(e, 2), e);
    // hard part
    F12 mz = conj(pow_z(e));                    // e^(-z)
    F12 m2z = cyclotomic_sqr(mz);               // e^(-2z)
    F12 m6z = mul(cyclotomic_sqr(m2z), m2z);    // e^(-6z)
    F12 p6zz = conj(pow_z(m6z));                // e^(6z^2)
    F12 p12zzz = pow_z(cyclotomic_sqr(p6zz));   // e^(12z^3)
    F12 k = mul(mul(p12zzz, p6zz), conj(m6z));  // e^(12z^3 + 6z^2 + 6z)
    F12 l = mul(k, m2z);                        // e^(12z^3 + 6z^2 + 4z)
    F12 n = mul(mul(k, p6zz), e);               // e^(12z^3 + 12z^2 + 6z + 1)
    F12 r = mul(frobenius(l, 1), n);
    r = mul(frobenius(k, 2), r);
    return mul(frobenius(mul(conj(e), l), 3), r);
}

static inline F12 pairing(const G1A& P, const G2A& Q, bool* ok) {
    F12 f = miller_product(&P, &Q, 1, ok);
    return *ok ? final_exponentiation(f) : F12::one();
}

}  // namespace pairing
}  // namespace ws
