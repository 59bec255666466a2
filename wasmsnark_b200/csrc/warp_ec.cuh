// warp_ec.cuh -- warp-cooperative forms of the latency-critical single-point chains (the Horner pass
// over MSM windows, the prover's final scalar multiplications; the reference's g?m_double / g?m_add chains,
// /root/reference/src/build_curve_jacobian_a0.js:174-385, in build_timesscalar.js:20-77).  One warp carries one
// point, state replicated on every lane; the independent products of each level of the XYZZ doubling / addition
// run on lanes 0..3 in ONE pass of the instruction stream and are exchanged with broadcast shuffles.  A warp
// instruction occupies the multiplier pipe for the same time whether 1 or 32 lanes are live, so this cuts a
// doubling from 9 sequential Montgomery products to 3 levels and an addition from 14 to 4.
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
__device__ __forceinline__ F lane_pick4(int lane, const F& a0, const F& a1, const F& a2, const F& a3) {   // lane 0 / 1 / 2 / other
    return select(lane == 0, a0, select(lane == 1, a1, select(lane == 2, a2, a3)));
}

// 2p [dbl-2008-s-1; the a ZZ^2 term joins level 1 when a != 0]: 3 product levels instead of 9-10 sequential products.
// All lanes: identical p in, identical 2p out.
template <class C>
__device__ __forceinline__ XYZZ<C> pdbl_warp4(const XYZZ<C>& p) {
    using F = typename C::F;
    if (p.is_inf()) return p;
    const int lane = threadIdx.x & 31;
    F U = dbl(p.y);
    // level 1: V = U^2 | XX = X^2 | (a != 0) ZZ^2
    F r1 = sqr(lane_pick4<F>(lane, U, p.x, p.zz, p.zz));
    F V = shfl_bcast_struct(r1, 0), XX = shfl_bcast_struct(r1, 1);
    F M = add(dbl(XX), XX);
    if (!C::A_ZERO) M = add(M, C::mul_a(shfl_bcast_struct(r1, 2)));
    // level 2: W = U V | S = X V | MM = M^2 | ZZ3 = V ZZ
    F r2 = mul(lane_pick4<F>(lane, U, p.x, M, V), lane_pick4<F>(lane, V, V, M, p.zz));
    F W = shfl_bcast_struct(r2, 0), S = shfl_bcast_struct(r2, 1), MM = shfl_bcast_struct(r2, 2), ZZ3 = shfl_bcast_struct(r2, 3);
    F X3 = sub(MM, dbl(S));
    // level 3: M (S - X3) | W Y | W ZZZ
    F r3 = mul(lane_pick4<F>(lane, M, W, W, W), lane_pick4<F>(lane, sub(S, X3), p.y, p.zzz, p.zzz));
    F t0 = shfl_bcast_struct(r3, 0), t1 = shfl_bcast_struct(r3, 1), ZZZ3 = shfl_bcast_struct(r3, 2);
    return XYZZ<C>{X3, sub(t0, t1), ZZ3, ZZZ3};
}

// a + b [add-2008-s]: 4 product levels instead of 14 sequential products; the rare cases (an operand at infinity,
// a == +-b) take the sequential path, uniformly on all lanes.
template <class C>
__device__ __forceinline__ XYZZ<C> padd_warp4(const XYZZ<C>& a, const XYZZ<C>& b) {
    using F = typename C::F;
    if (a.is_inf()) return b;
    if (b.is_inf()) return a;
    const int lane = threadIdx.x & 31;
    // level 1: U1 = X1 ZZ2 | U2 = X2 ZZ1 | S1 = Y1 ZZZ2 | S2 = Y2 ZZZ1
    F r1 = mul(lane_pick4<F>(lane, a.x, b.x, a.y, b.y), lane_pick4<F>(lane, b.zz, a.zz, b.zzz, a.zzz));
    F U1 = shfl_bcast_struct(r1, 0), U2 = shfl_bcast_struct(r1, 1), S1 = shfl_bcast_struct(r1, 2), S2 = shfl_bcast_struct(r1, 3);
    F P = sub(U2, U1), R = sub(S2, S1);
    if (P.is_zero()) return R.is_zero() ? pdbl_call<C>(a) : XYZZ<C>::inf();
    // level 2: PP = P^2 | RR = R^2 | ZZ12 = ZZ1 ZZ2 | ZZZ12 = ZZZ1 ZZZ2
    F r2 = mul(lane_pick4<F>(lane, P, R, a.zz, a.zzz), lane_pick4<F>(lane, P, R, b.zz, b.zzz));
    F PP = shfl_bcast_struct(r2, 0), RR = shfl_bcast_struct(r2, 1), ZZ12 = shfl_bcast_struct(r2, 2), ZZZ12 = shfl_bcast_struct(r2, 3);
    // level 3: PPP = P PP | Q = U1 PP | ZZ3 = ZZ12 PP
    F r3 = mul(lane_pick4<F>(lane, P, U1, ZZ12, ZZ12), PP);
    F PPP = shfl_bcast_struct(r3, 0), Q = shfl_bcast_struct(r3, 1), ZZ3 = shfl_bcast_struct(r3, 2);
    F X3 = sub(sub(RR, PPP), dbl(Q));
    // level 4: R (Q - X3) | S1 PPP | ZZZ12 PPP
    F r4 = mul(lane_pick4<F>(lane, R, S1, ZZZ12, ZZZ12), lane_pick4<F>(lane, sub(Q, X3), PPP, PPP, PPP));
    F t0 = shfl_bcast_struct(r4, 0), t1 = shfl_bcast_struct(r4, 1), ZZZ3 = shfl_bcast_struct(r4, 2);
    return XYZZ<C>{X3, sub(t0, t1), ZZ3, ZZZ3};
}

// 2^k * P, warp-cooperative.
template <class C>
__device__ __forceinline__ XYZZ<C> pdbl_times_warp(XYZZ<C> r, int k) {
    if (r.is_inf()) return r;
#pragma unroll 1
    for (int d = 0; d < k; d++) r = pdbl_warp4<C>(r);
    return r;
}

// k * P (4-bit fixed window), all lanes carrying the same state; same element as g?m_timesScalar.
// The table lives across the lanes: lane i works out i * P by double-and-add on its own (lanes 0..15 matter), the
// window's entry is fetched with a broadcast shuffle.
template <class C>
__device__ __forceinline__ XYZZ<C> times_scalar_w4_warp(const XYZZ<C>& p, const uint8_t* k, int len) {
    const int lane = threadIdx.x & 31;
    XYZZ<C> mine = XYZZ<C>::inf();
#pragma unroll 1
    for (int b = 3; b >= 0; b--) {
        mine = pdbl_call<C>(mine);
        if ((lane >> b) & 1) mine = padd_call<C>(mine, p);
    }
    XYZZ<C> r = XYZZ<C>::inf();
#pragma unroll 1
    for (int i = 2 * len - 1; i >= 0; i--) {
        uint32_t nib = (k[i >> 1] >> ((i & 1) * 4)) & 15u;
        r = pdbl_times_warp<C>(r, 4);
        if (nib) r = padd_warp4<C>(r, shfl_bcast_struct(mine, (int)nib));
    }
    return r;
}

// Horner over the MSM windows, executed by one full warp.
template <class C>
__device__ __forceinline__ XYZZ<C> msm_horner_windows_warp(const MsmPlanLite& p, const XYZZ<C>* R) {
    XYZZ<C> res = R[p.nwin - 1];
    for (int w = p.nwin - 2; w >= 0; w--) {
        res = pdbl_times_warp<C>(res, p.c);
        res = padd_warp4<C>(res, R[w]);
    }
    return res;
}
#endif

}  // namespace ws
