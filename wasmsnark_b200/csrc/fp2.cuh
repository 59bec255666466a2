// fp2.cuh -- quadratic extension Fp2 = Fp[u]/(u^2 + 1) for BN128 G2.
//
// Replaces the reference's generated f2m module (non-residue = -1, mul-by-NR = f1m_neg:
// /root/reference/src/bn128/build_bn128.js:40):
//   f2m_add/sub/neg     /root/reference/src/build_f2m.js:30-108
//   f2m_mul             /root/reference/src/build_f2m.js:127-163   (Karatsuba, 3 base muls)
//   f2m_square          /root/reference/src/build_f2m.js:186-228   (2 base muls)
//   f2m_inverse         /root/reference/src/build_f2m.js:275-333
//   f2m_isZero          /root/reference/src/build_f2m.js:335-351
// Element layout (c0, c1), each a canonical Montgomery base-field element.
#pragma once
#include "fp.cuh"

namespace ws {

template <class P>
struct Fp2T {
    using Base = Fp<P>;
    Base c0, c1;
    HD static Fp2T zero() { return Fp2T{Base::zero(), Base::zero()}; }
    HD static Fp2T one() { return Fp2T{Base::one(), Base::zero()}; }
    HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    HD bool operator==(const Fp2T& o) const { return c0 == o.c0 && c1 == o.c1; }
    HD bool operator!=(const Fp2T& o) const { return !(*this == o); }
};

template <class P>
HD Fp2T<P> add(const Fp2T<P>& a, const Fp2T<P>& b) { return Fp2T<P>{add(a.c0, b.c0), add(a.c1, b.c1)}; }
template <class P>
HD Fp2T<P> sub(const Fp2T<P>& a, const Fp2T<P>& b) { return Fp2T<P>{sub(a.c0, b.c0), sub(a.c1, b.c1)}; }
template <class P>
HD Fp2T<P> neg(const Fp2T<P>& a) { return Fp2T<P>{neg(a.c0), neg(a.c1)}; }
template <class P>
HD Fp2T<P> dbl(const Fp2T<P>& a) { return Fp2T<P>{dbl(a.c0), dbl(a.c1)}; }

template <class P>
HD Fp2T<P> mul(const Fp2T<P>& x, const Fp2T<P>& y) {
    Fp<P> A = mul_ool(x.c0, y.c0);
    Fp<P> B = mul_ool(x.c1, y.c1);
    Fp<P> C = mul_ool(add(x.c0, x.c1), add(y.c0, y.c1));
    return Fp2T<P>{sub(A, B), sub(C, add(A, B))};
}

template <class P>
HD Fp2T<P> sqr(const Fp2T<P>& x) {
    // (x0 + x1)(x0 - x1), 2 x0 x1
    Fp<P> AB = mul_ool(x.c0, x.c1);
    Fp<P> r0 = mul_ool(add(x.c0, x.c1), sub(x.c0, x.c1));
    return Fp2T<P>{r0, dbl(AB)};
}

template <class P>
HD Fp2T<P> inv(const Fp2T<P>& x) {
    Fp<P> t = inv(add(sqr(x.c0), sqr(x.c1)));
    return Fp2T<P>{mul_ool(x.c0, t), neg(mul_ool(x.c1, t))};
}

template <class P>
HD Fp2T<P> select(bool c, const Fp2T<P>& a, const Fp2T<P>& b) {
    return Fp2T<P>{select(c, a.c0, b.c0), select(c, a.c1, b.c1)};
}

template <class P>
HD Fp2T<P> from_mont(const Fp2T<P>& a) { return Fp2T<P>{from_mont(a.c0), from_mont(a.c1)}; }

using Fq2 = Fp2T<BN128_Fq>;

}  // namespace ws
