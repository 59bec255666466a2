// ec.cuh -- short-Weierstrass group law for the MSM / prover kernels.
//
// Replaces the reference's generated curve modules
//   BN128 G1/G2 (a = 0, Jacobian)       /root/reference/src/build_curve_jacobian_a0.js
//       isZero :33-47  zero :87-113  double :174-235  add :280-385  neg :387-406  affine :424-457
//   MNT6753 G1/G2 (a != 0, projective)  /root/reference/src/build_curve_projective.js:202-516,624-652
// and the point-loading rule of the reference's multiexp
//   x == 0  =>  the affine key point is the point at infinity
//                                        /root/reference/src/build_multiexp.js:335-349
//
// Internal representation: XYZZ (X, Y, ZZ, ZZZ) with x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2;
// infinity <=> ZZ == 0.  Mixed addition costs 8M+2S against 7M+4S for Jacobian and the
// formulas are complete here because P = +-Q and infinity are tested explicitly, exactly
// the cases the reference's add handles (:322-356).  The group element computed is the
// same as the reference's; byte-level equality is established after affine normalisation
// (to_affine below == g1m_affine), the way the reference's own test compares
// (/root/reference/test/bn128_prover.js:31-46).
#pragma once
#include "fp.cuh"
#include "fp2.cuh"

namespace ws {

// Curve traits: field type + the coefficient a (b never enters the formulas).
struct G1Curve {
    using F = Fq;
    static constexpr bool A_ZERO = true;
    static constexpr bool PROJECTIVE = false;   // external triple is Jacobian (x = X/Z^2, y = Y/Z^3)
    static constexpr int SCALAR_LIMBS = 8;
    HD static F mul_a(const F&) { return F::zero(); }
};
struct G2Curve {
    using F = Fq2;
    static constexpr bool A_ZERO = true;
    static constexpr bool PROJECTIVE = false;
    static constexpr int SCALAR_LIMBS = 8;
    HD static F mul_a(const F&) { return F::zero(); }
};

// MNT6753 G1: y^2 = x^3 + 11 x + b over the 753-bit field (24 limbs); the reference builds it with
// homogeneous projective coordinates (x = X/Z, y = Y/Z) and mulByA = multiplication by 11
// (/root/reference/src/mnt6753/build_mnt6753.js:90-109).
using MntFq = Fp<MNT6753_Fq>;
struct MntG1Curve {
    using F = MntFq;
    static constexpr bool A_ZERO = false;
    static constexpr bool PROJECTIVE = true;
    static constexpr int SCALAR_LIMBS = 24;
    HD static F mul_a(const F& x) {          // 11 x = 8x + 2x + x
        F x2 = dbl(x), x8 = dbl(dbl(x2));
        return add(add(x8, x2), x);
    }
};

// BLS12-381: G1 y^2 = x^3 + 4 over the 381-bit field (12 limbs), G2 y^2 = x^3 + 4(1 + u) over Fq2 = Fq[u]/(u^2 + 1);
// both built by the reference with the a = 0 Jacobian curve module and 32-byte scalars
// (/root/reference/src/bls12381/build_bls12381.js:33-43).
using BlsFq = Fp<BLS12_381_Fq>;
using BlsFr = Fp<BLS12_381_Fr>;
using BlsFq2 = Fp2T<BLS12_381_Fq>;
struct BlsG1Curve {
    using F = BlsFq;
    static constexpr bool A_ZERO = true;
    static constexpr bool PROJECTIVE = false;
    static constexpr int SCALAR_LIMBS = 8;
    HD static F mul_a(const F&) { return F::zero(); }
};
struct BlsG2Curve {
    using F = BlsFq2;
    static constexpr bool A_ZERO = true;
    static constexpr bool PROJECTIVE = false;
    static constexpr int SCALAR_LIMBS = 8;
    HD static F mul_a(const F&) { return F::zero(); }
};

template <class C>
struct Affine {
    using F = typename C::F;
    F x, y;
    HD bool is_inf() const { return x.is_zero(); }   // reference's loading rule
};

template <class C>
struct XYZZ {
    using F = typename C::F;
    F x, y, zz, zzz;
    HD static XYZZ inf() { return XYZZ{F::zero(), F::one(), F::zero(), F::zero()}; }
    HD bool is_inf() const { return zz.is_zero(); }
    HD static XYZZ from_affine(const Affine<C>& p) {
        if (p.is_inf()) return inf();
        return XYZZ{p.x, p.y, F::one(), F::one()};
    }
};

// Jacobian triple, only as the external result format (g1_multiexp returns X|Y|Z).
template <class C>
struct Jac {
    using F = typename C::F;
    F x, y, z;
};

template <class C>
HD XYZZ<C> pneg(const XYZZ<C>& p) { return XYZZ<C>{p.x, neg(p.y), p.zz, p.zzz}; }

// 2*(x, y) for an affine, finite point   [mdbl-2008-s-1].  Out of line: it is the rare
// P == Q branch of the mixed addition and would only bloat the accumulate kernel.
template <class C>
HD_NOINLINE XYZZ<C> dbl_affine(const Affine<C>& p) {
    using F = typename C::F;
    F U = dbl(p.y);
    F V = sqr(U);
    F W = mul(U, V);
    F S = mul(p.x, V);
    F XX = sqr(p.x);
    F M = add(dbl(XX), XX);
    if (!C::A_ZERO) M = add(M, C::mul_a(F::one()));
    F X3 = sub(sqr(M), dbl(S));
    F Y3 = sub(mul(M, sub(S, X3)), mul(W, p.y));
    return XYZZ<C>{X3, Y3, V, W};
}

// 2*P   [dbl-2008-s-1]
template <class C>
HD XYZZ<C> pdbl(const XYZZ<C>& p) {
    using F = typename C::F;
    if (p.is_inf()) return p;
    F U = dbl(p.y);
    F V = sqr(U);
    F W = mul(U, V);
    F S = mul(p.x, V);
    F XX = sqr(p.x);
    F M = add(dbl(XX), XX);
    if (!C::A_ZERO) M = add(M, C::mul_a(sqr(p.zz)));
    F X3 = sub(sqr(M), dbl(S));
    F Y3 = sub(mul(M, sub(S, X3)), mul(W, p.y));
    return XYZZ<C>{X3, Y3, mul(V, p.zz), mul(W, p.zzz)};
}

// acc + (x2, y2), affine finite q   [madd-2008-s]; q infinite must be filtered by the caller
template <class C>
HD XYZZ<C> pmadd(const XYZZ<C>& a, const Affine<C>& q) {
    using F = typename C::F;
    if (a.is_inf()) return XYZZ<C>{q.x, q.y, F::one(), F::one()};
    F U2 = mul(q.x, a.zz);
    F S2 = mul(q.y, a.zzz);
    F P = sub(U2, a.x);
    F R = sub(S2, a.y);
    if (P.is_zero()) {
        if (R.is_zero()) return dbl_affine<C>(q);
        return XYZZ<C>::inf();
    }
    F PP = sqr(P);
    F PPP = mul(P, PP);
    F Q = mul(a.x, PP);
    F X3 = sub(sub(sqr(R), PPP), dbl(Q));
    F Y3 = sub(mul(R, sub(Q, X3)), mul(a.y, PPP));
    return XYZZ<C>{X3, Y3, mul(a.zz, PP), mul(a.zzz, PPP)};
}

// a + b   [add-2008-s]
template <class C>
HD XYZZ<C> padd(const XYZZ<C>& a, const XYZZ<C>& b) {
    using F = typename C::F;
    if (a.is_inf()) return b;
    if (b.is_inf()) return a;
    F U1 = mul(a.x, b.zz);
    F U2 = mul(b.x, a.zz);
    F S1 = mul(a.y, b.zzz);
    F S2 = mul(b.y, a.zzz);
    F P = sub(U2, U1);
    F R = sub(S2, S1);
    if (P.is_zero()) {
        if (R.is_zero()) return pdbl<C>(a);
        return XYZZ<C>::inf();
    }
    F PP = sqr(P);
    F PPP = mul(P, PP);
    F Q = mul(U1, PP);
    F X3 = sub(sub(sqr(R), PPP), dbl(Q));
    F Y3 = sub(mul(R, sub(Q, X3)), mul(S1, PPP));
    return XYZZ<C>{X3, Y3, mul(mul(a.zz, b.zz), PP), mul(mul(a.zzz, b.zzz), PPP)};
}

// g?m_affine: (x, y, 1), infinity -> (0, 1, 0)   (build_curve_jacobian_a0.js:424-457)
template <class C>
HD Jac<C> to_affine(const XYZZ<C>& p) {
    using F = typename C::F;
    if (p.is_inf()) return Jac<C>{F::zero(), F::one(), F::zero()};
    F i = inv(mul(p.zz, p.zzz));
    F izz = mul(i, p.zzz);
    F izzz = mul(i, p.zz);
    return Jac<C>{mul(p.x, izz), mul(p.y, izzz), F::one()};
}

// A Jacobian triple (X, Y, Z) read as XYZZ: (X, Y, Z^2, Z^3).
template <class C>
HD XYZZ<C> from_jac(const Jac<C>& p) {
    using F = typename C::F;
    if (p.z.is_zero()) return XYZZ<C>::inf();
    F zz = sqr(p.z);
    return XYZZ<C>{p.x, p.y, zz, mul(zz, p.z)};
}

// XYZZ -> Jacobian triple of the same point: (X*ZZ, Y*ZZZ, ZZ)  (Z := ZZ, so Z^2 = ZZ^2, Z^3 = ZZ^3 = ZZZ^2)
template <class C>
HD Jac<C> to_jac(const XYZZ<C>& p) {
    using F = typename C::F;
    if (p.is_inf()) return Jac<C>{F::zero(), F::one(), F::zero()};
    return Jac<C>{mul(p.x, p.zz), mul(p.y, p.zzz), p.zz};
}

// External triple <-> XYZZ for whichever convention the curve's reference module uses.
template <class C>
HD Jac<C> to_ext(const XYZZ<C>& p) {
    using F = typename C::F;
    if (!C::PROJECTIVE) return to_jac<C>(p);
    if (p.is_inf()) return Jac<C>{F::zero(), F::one(), F::zero()};
    return Jac<C>{mul(p.x, p.zzz), mul(p.y, p.zz), mul(p.zz, p.zzz)};     // (X ZZZ : Y ZZ : ZZ ZZZ)
}
template <class C>
HD XYZZ<C> from_ext(const Jac<C>& p);

// Jacobian doubling for a = 0 [dbl-2009-l], the same formula as the reference's g?m_double
// (build_curve_jacobian_a0.js:174-235): 2M + 5S in three dependency levels of three independent
// products each, which is what the single-thread Horner tail wants.
template <class C>
HD Jac<C> jdbl_a0(const Jac<C>& p) {
    using F = typename C::F;
    if (p.z.is_zero()) return p;
    F A = sqr(p.x);
    F B = sqr(p.y);
    F G = mul(p.y, p.z);
    F Cc = sqr(B);
    F D = sqr(add(p.x, B));
    F E = add(dbl(A), A);
    F Fv = sqr(E);
    D = dbl(sub(sub(D, A), Cc));
    F X3 = sub(Fv, dbl(D));
    F C8 = dbl(dbl(dbl(Cc)));
    F Y3 = sub(mul(sub(D, X3), E), C8);
    return Jac<C>{X3, Y3, dbl(G)};
}

// Out-of-line forms for the latency-bound tail kernels (segment reduce, bit sums, Horner):
// one copy of the group law per curve instead of one per call site keeps the kernels small.
template <class C>
HD_NOINLINE void padd_to(XYZZ<C>* r, const XYZZ<C>* a, const XYZZ<C>* b) { *r = padd<C>(*a, *b); }
template <class C>
HD_NOINLINE void pdbl_to(XYZZ<C>* r, const XYZZ<C>* a) { *r = pdbl<C>(*a); }
template <class C>
HD_NOINLINE void pmadd_to(XYZZ<C>* r, const XYZZ<C>* a, const Affine<C>* q) { *r = pmadd<C>(*a, *q); }
template <class C>
HD XYZZ<C> pmadd_call(const XYZZ<C>& a, const Affine<C>& q) {
    XYZZ<C> r;
    pmadd_to<C>(&r, &a, &q);
    return r;
}
template <class C>
HD XYZZ<C> padd_call(const XYZZ<C>& a, const XYZZ<C>& b) {
    XYZZ<C> r;
    padd_to<C>(&r, &a, &b);
    return r;
}
// Throughput-bound tail kernel (segment reduction): inline for the 8-limb base field (registers instead of the
// local-memory round trip of a call), out of line for extension / wide fields to keep the kernel small.
template <class C>
HD XYZZ<C> padd_seg(const XYZZ<C>& a, const XYZZ<C>& b) {
    if constexpr (sizeof(typename C::F) == 32) return padd<C>(a, b);
    else return padd_call<C>(a, b);
}
template <class C>
HD XYZZ<C> pdbl_call(const XYZZ<C>& a) {
    XYZZ<C> r;
    pdbl_to<C>(&r, &a);
    return r;
}

template <class C>
HD XYZZ<C> from_ext(const Jac<C>& p) {
    using F = typename C::F;
    if (!C::PROJECTIVE) return from_jac<C>(p);
    if (p.z.is_zero()) return XYZZ<C>::inf();
    F zz = sqr(p.z);
    return XYZZ<C>{mul(p.x, p.z), mul(p.y, zz), zz, mul(zz, p.z)};       // x = XZ/Z^2, y = YZ^2/Z^3
}

// k * P, MSB first double-and-add over a little-endian byte string, as
// g?m_timesScalar (/root/reference/src/build_timesscalar.js:20-77).
template <class C>
HD XYZZ<C> times_scalar(const XYZZ<C>& p, const uint8_t* k, int len) {
    XYZZ<C> r = XYZZ<C>::inf();
    for (int i = len - 1; i >= 0; i--) {
        uint32_t byte = k[i];
        for (int b = 7; b >= 0; b--) {
            r = pdbl_call<C>(r);
            if ((byte >> b) & 1) r = padd_call<C>(r, p);
        }
    }
    return r;
}

using G1Affine = Affine<G1Curve>;
using G2Affine = Affine<G2Curve>;
using G1X = XYZZ<G1Curve>;
using G2X = XYZZ<G2Curve>;

}  // namespace ws
