// msm_impl.cuh -- host-side driver of the MSM pipeline, instantiated per curve in
// msm_g1.cu / msm_g2.cu (separate translation units so they compile in parallel).
#pragma once
#include <cuda_runtime.h>
#include "../../include/wasmsnark_b200.h"
#include "ctx.cuh"
#include "msm.cuh"
#include "msm_api.cuh"

#define WS_API __attribute__((visibility("default")))

namespace ws {

// d_out: Jacobian triple (may be null); d_out_xyzz: XYZZ sum for device consumers (may be null)
template <class C>
inline int multiexp_dev(ws_ctx* c, const void* d_scalars, const void* d_points, size_t n, void* d_out, int window_bits, cudaStream_t st, void* d_out_xyzz = nullptr, bool own_timer = true, int arena_slot = 1, int table_c = 0) {
    if (n > 0x7FFFFFFFull) return c->fail(WS_E_ARG, "n too large");
    if (((uintptr_t)d_scalars | (uintptr_t)d_points | (uintptr_t)d_out | (uintptr_t)d_out_xyzz) & 15) return c->fail(WS_E_ARG, "device pointers must be 16-byte aligned");
    int cb;
    uint32_t T;
    constexpr int SBITS = 32 * C::SCALAR_LIMBS;
    msm_choose((uint32_t)n, &cb, &T, SBITS, table_c != 0);
    if (table_c) {            // d_points is a window table built for exactly this window size
        cb = table_c;
    } else if (window_bits) {
        if (window_bits < 2 || window_bits > 16) return c->fail(WS_E_ARG, "window_bits must be in 2..16");
        cb = window_bits;
    }
    if (c->force_T) T = c->force_T;
    MsmPlan p = msm_make_plan((uint32_t)n, cb, T, c->seg_len, SBITS, table_c != 0);
    if ((uint64_t)p.ndig * n >= 0xFFFFFF00ull) return c->fail(WS_E_ARG, "n too large: digit windows x n must fit 32-bit offsets");
    size_t need = msm_carve<C>(p, nullptr, nullptr);
    uint8_t* base = (uint8_t*)c->arena(arena_slot, need);
    if (!base) return c->fail_cuda("msm workspace");
    MsmWorkspace w;
    msm_carve<C>(p, base, &w);
    // a public entry (own_timer): chain after the previous asynchronous call on this context (shared arenas)
    struct Chain {
        ws_ctx* c;
        cudaStream_t st;
        bool on;
        Chain(ws_ctx* c_, cudaStream_t st_, bool on_) : c(c_), st(st_), on(on_) {
            if (on) c->enter_stream(st);
        }
        ~Chain() {
            if (on) c->leave_stream(st);
        }
    } chain(c, st, own_timer);
    if (own_timer) c->timer.begin(st);
    cudaError_t e = msm_launch<C>(p, w, (const uint32_t*)d_scalars, (const Affine<C>*)d_points, (XYZZ<C>*)d_out_xyzz, (Jac<C>*)d_out, st, own_timer ? &c->timer : nullptr, &c->launches, c->acc_variant);
    if (e != cudaSuccess) return c->fail_cuda("msm launch");
    return WS_OK;
}

template <class C>
inline int multiexp_host(ws_ctx* c, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out) {
    WsLock ws_lock_(c);
    if (!c || !out || (n && (!scalars || !points))) return c ? c->fail(WS_E_ARG, "null buffer") : WS_E_ARG;
    cudaSetDevice(c->device);
    constexpr size_t PS = sizeof(Affine<C>), OS = sizeof(Jac<C>), SS = 4 * C::SCALAR_LIMBS;
    uint8_t* d = (uint8_t*)c->arena(0, align256(n * SS) + align256(n * PS) + 512);
    if (!d) return c->fail_cuda("input workspace");
    uint8_t *ds = d, *dp = d + align256(n * SS), *dout = dp + align256(n * PS);
    if (n) {
        if (cudaMemcpyAsync(ds, scalars, n * SS, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaMemcpyAsync(dp, points, n * PS, cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
            return c->fail_cuda("multiexp: host-to-device copy");
    }
    int rc = multiexp_dev<C>(c, ds, dp, n, dout, 0, c->stream);
    if (rc) return rc;
    if (cudaMemcpyAsync(out, dout, OS, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return c->fail_cuda("multiexp: device-to-host copy");
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return c->fail_cuda("multiexp");
    return WS_OK;
}





// Builds the window table of `n` points for window size table_c (see k_msm_build_table); d_table must hold
// msm_table_windows(table_c) * n affine points.
inline int msm_table_windows(int table_c) { return (257 + table_c - 1) / table_c; }

template <class C>
inline int build_table_dev(ws_ctx* c, const void* d_points, size_t n, int table_c, void* d_table, cudaStream_t st) {
    if (table_c < 8 || table_c > 22) return c->fail(WS_E_ARG, "table window bits must be in 8..22");
    if (n == 0) return WS_OK;
    int W = msm_table_windows(table_c);
    if ((uint64_t)W * n >= 0x80000000ull) return c->fail(WS_E_ARG, "table too large for 31-bit point indices");
    c->launches += 1;
    k_msm_build_table<C><<<(unsigned)((n + 63) / 64), 64, 0, st>>>((const Affine<C>*)d_points, (uint32_t)n, table_c, W, (Affine<C>*)d_table);
    return cudaGetLastError() == cudaSuccess ? WS_OK : c->fail_cuda("build_table");
}

}  // namespace ws
