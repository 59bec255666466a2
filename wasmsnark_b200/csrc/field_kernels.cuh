// field_kernels.cuh -- element-wise batch kernels over Montgomery field arrays (config C1/C5):
// the device form of f1m_mul / add / sub / square / toMontgomery / fromMontgomery / inverse
// (/root/reference/src/build_f1m.js:67-101, :235-436, :439-783) and of the batch loops
// fft_toMontgomeryN / fromMontgomeryN / mulN (/root/reference/src/build_fft.js:418-547).
// One element per thread, 128-bit global loads/stores, whole number in registers.
#pragma once
#include <cuda_runtime.h>
#include "fp.cuh"
#include "fp2.cuh"
#include "../../include/wasmsnark_b200.h"

namespace ws {

template <class T>
__device__ __forceinline__ T ld_vec(const T* p) {
    static_assert(sizeof(T) % 16 == 0, "");
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
    return r;
}
template <class T>
__device__ __forceinline__ void st_vec(T* p, const T& v) {
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
}

template <class P, int OP>
static __global__ void __launch_bounds__(128) k_fp_batch(const Fp<P>* __restrict__ x, const Fp<P>* __restrict__ y, Fp<P>* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp<P> a = ld_vec(x + i), r;
    if (OP == WS_OP_MUL) r = mul(a, ld_vec(y + i));
    else if (OP == WS_OP_ADD) r = add(a, ld_vec(y + i));
    else if (OP == WS_OP_SUB) r = sub(a, ld_vec(y + i));
    else if (OP == WS_OP_SQUARE) r = sqr(a);
    else if (OP == WS_OP_TOMONT) r = to_mont(a);
    else if (OP == WS_OP_FROMMONT) r = from_mont(a);
    else r = inv(a);
    st_vec(out + i, r);
}

template <class P>
inline void launch_fp_batch(int op, const void* x, const void* y, void* out, size_t n, cudaStream_t st) {
    unsigned g = (unsigned)((n + 127) / 128);
    const Fp<P>* X = (const Fp<P>*)x;
    const Fp<P>* Y = (const Fp<P>*)y;
    Fp<P>* O = (Fp<P>*)out;
    switch (op) {
        case WS_OP_MUL: k_fp_batch<P, WS_OP_MUL><<<g, 128, 0, st>>>(X, Y, O, n); break;
        case WS_OP_ADD: k_fp_batch<P, WS_OP_ADD><<<g, 128, 0, st>>>(X, Y, O, n); break;
        case WS_OP_SUB: k_fp_batch<P, WS_OP_SUB><<<g, 128, 0, st>>>(X, Y, O, n); break;
        case WS_OP_SQUARE: k_fp_batch<P, WS_OP_SQUARE><<<g, 128, 0, st>>>(X, Y, O, n); break;
        case WS_OP_TOMONT: k_fp_batch<P, WS_OP_TOMONT><<<g, 128, 0, st>>>(X, Y, O, n); break;
        case WS_OP_FROMMONT: k_fp_batch<P, WS_OP_FROMMONT><<<g, 128, 0, st>>>(X, Y, O, n); break;
        default: k_fp_batch<P, WS_OP_INVERSE><<<g, 128, 0, st>>>(X, Y, O, n); break;
    }
}

static __global__ void __launch_bounds__(128) k_fq2_mul(const Fq2* __restrict__ x, const Fq2* __restrict__ y, Fq2* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_vec(out + i, mul(ld_vec(x + i), ld_vec(y + i)));
}

}  // namespace ws
