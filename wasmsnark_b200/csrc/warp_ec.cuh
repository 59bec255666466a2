// warp_ec.cuh -- warp-cooperative forms of the latency-critical single-point chains (the Horner pass
// over MSM windows, the prover's final scalar multiplications).  One warp carries one point, state
// replicated on every lane; the three independent products of each level of the Jacobian doubling
// (dbl-2009-l, a = 0: /root/reference/src/build_curve_jacobian_a0.js:174-235 computes the same formula)
// run on lanes 0, 1, 2 in ONE pass of the instruction stream and are exchanged with broadcast
// shuffles.  A warp instruction occupies the multiplier pipe for the same time whether 1 or 32 lanes
// are live, so this cuts a doubling from 7 sequential Montgomery products to 3.
#pragma once
#include "ec.cuh"

namespace ws {

struct MsmPlanLite {
    int nwin, c;
};

#ifdef __CUDACC__
template <class T>
__device__ __forceinline__ T shfl_bcast_struct(const T& v, int src) {
    static_assert(sizeof(T) % 4 == 0, "");
    T r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); i++) d[i] = __shfl_sync(0xFFFFFFFFu, s[i], src);
    return r;
}

template <class F>
__device__ __forceinline__ F lane_pick(int lane, const F& a0, const F& a1, const F& a2) {   // lane 0 / 1 / other
    return select(lane == 0, a0, select(lane == 1, a1, a2));
}

// All 32 lanes must call this with identical p; returns 2p, identical on all lanes.
template <class C>
__device__ __forceinline__ Jac<C> jdbl_a0_warp3(const Jac<C>& p) {
    using F = typename C::F;
    if (p.z.is_zero()) return p;
    const int lane = threadIdx.x & 31;
    // level 1: A = X^2 | B = Y^2 | G = Y Z
    F r1 = mul(lane_pick<F>(lane, p.x, p.y, p.y), lane_pick<F>(lane, p.x, p.y, p.z));
    F A = shfl_bcast_struct(r1, 0), B = shfl_bcast_struct(r1, 1), G = shfl_bcast_struct(r1, 2);
    F E = add(dbl(A), A);
    // level 2: D' = (X + B)^2 | C = B^2 | F = E^2
    F r2 = sqr(lane_pick<F>(lane, add(p.x, B), B, E));
    F Dp = shfl_bcast_struct(r2, 0), Cc = shfl_bcast_struct(r2, 1), Fv = shfl_bcast_struct(r2, 2);
    // level 3 (every lane, same operands)
    F D = dbl(sub(sub(Dp, A), Cc));
    F X3 = sub(Fv, dbl(D));
    F Y3 = sub(mul(sub(D, X3), E), dbl(dbl(dbl(Cc))));
    return Jac<C>{X3, Y3, dbl(G)};
}

// 2^k * P, warp-cooperative.
template <class C>
__device__ __forceinline__ XYZZ<C> pdbl_times_warp(XYZZ<C> r, int k) {
    if (r.is_inf()) return r;
    if (C::A_ZERO) {
        Jac<C> j = to_jac<C>(r);
#pragma unroll 1
        for (int d = 0; d < k; d++) j = jdbl_a0_warp3<C>(j);
        return from_jac<C>(j);
    }
    for (int d = 0; d < k; d++) r = pdbl_call<C>(r);
    return r;
}

// k * P (4-bit fixed window), all lanes carrying the same state; same element as g?m_timesScalar.
template <class C>
__device__ __forceinline__ XYZZ<C> times_scalar_w4_warp(const XYZZ<C>& p, const uint8_t* k, int len) {
    XYZZ<C> tab[16];
    tab[0] = XYZZ<C>::inf();
    tab[1] = p;
    for (int i = 2; i < 16; i++) tab[i] = (i & 1) ? padd_call<C>(tab[i - 1], p) : pdbl_call<C>(tab[i >> 1]);
    XYZZ<C> r = XYZZ<C>::inf();
    for (int i = 2 * len - 1; i >= 0; i--) {
        uint32_t nib = (k[i >> 1] >> ((i & 1) * 4)) & 15u;
        r = pdbl_times_warp<C>(r, 4);
        if (nib) r = padd_call<C>(r, tab[nib]);
    }
    return r;
}

// Horner over the MSM windows, executed by one full warp.
template <class C>
__device__ __forceinline__ XYZZ<C> msm_horner_windows_warp(const MsmPlanLite& p, const XYZZ<C>* R) {
    XYZZ<C> res = R[p.nwin - 1];
    for (int w = p.nwin - 2; w >= 0; w--) {
        res = pdbl_times_warp<C>(res, p.c);
        res = padd_call<C>(res, R[w]);
    }
    return res;
}
#endif

}  // namespace ws
