// msm.cuh -- CUDA kernels and launch sequence of the MSM pipeline (see msm_core.cuh for the
// algorithm and the reference citations).  sm_100a; integer pipe bound, no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include "ctx.cuh"
#include "msm_core.cuh"
#include "warp_ec.cuh"

namespace ws {

// ------------------------------------------------------------------ small utilities
template <class T>
__device__ __forceinline__ T shfl_down_struct(const T& v, int off) {
    static_assert(sizeof(T) % 4 == 0, "word-sized structs only");
    T r;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); i++) dst[i] = __shfl_down_sync(0xFFFFFFFFu, src[i], off);
    return r;
}

// 16-byte vector copy of a POD struct from global memory (structs are 32-byte multiples and
// the arrays are 256-byte aligned device allocations).
template <class T>
__device__ __forceinline__ T ldg_struct(const T* p) {
    static_assert(sizeof(T) % 16 == 0, "");
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = __ldg(s + i);
    return r;
}
template <class T>
__device__ __forceinline__ void stg_struct(T* p, const T& v) {
    static_assert(sizeof(T) % 16 == 0, "");
    const uint4* s = reinterpret_cast<const uint4*>(&v);
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
}

// ------------------------------------------------------------------ exclusive scan (u32)
// mode 0: value = in[i];  mode 1: value = ceil(in[i] / T).   out has n + 1 entries.
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t scan_value(uint32_t v, int mode, uint32_t T) { return mode ? (v + T - 1) / T : v; }

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                              uint32_t* __restrict__ tile_sums, uint32_t n, int mode, uint32_t T) {
    __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint32_t x = (base + k < n) ? scan_value(in[base + k], mode, T) : 0u;
        v[k] = sum;           // exclusive within the thread
        sum += x;
    }
    // inclusive warp scan of the thread sums
    uint32_t inc = sum;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t t = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 0u;
        uint32_t ti = t;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t u = __shfl_up_sync(0xFFFFFFFFu, ti, off);
            if (lane >= off) ti += u;
        }
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t;   // exclusive warp offsets
        if (lane == SCAN_THREADS / 32 - 1) tile_sums[blockIdx.x] = ti;
    }
    __syncthreads();
    uint32_t off0 = warp_tot[wid] + (inc - sum);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) out[base + k] = off0 + v[k];
}

// single block: exclusive scan of the tile sums in place; total -> out[n]
static __global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* __restrict__ tile_sums, uint32_t ntiles, uint32_t* __restrict__ out, uint32_t n) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t base = 0; base < ntiles; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t x = i < ntiles ? tile_sums[i] : 0u;
        uint32_t inc = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t t = warp_tot[lane], ti = t;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t u = __shfl_up_sync(0xFFFFFFFFu, ti, off);
                if (lane >= off) ti += u;
            }
            warp_tot[lane] = ti - t;
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t excl = carry + warp_tot[wid] + inc - x;
        if (i < ntiles) tile_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ tile_sums, uint32_t n) {
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t add = tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) out[base + k] += add;
}

// out[0..n] = exclusive scan; tile_sums: scratch of ceil(n / SCAN_TILE) words
inline void launch_scan(const uint32_t* in, uint32_t* out, uint32_t* tile_sums, uint32_t n, int mode, uint32_t T, cudaStream_t st) {
    uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tiles<<<ntiles, SCAN_THREADS, 0, st>>>(in, out, tile_sums, n, mode, T);
    k_scan_sums<<<1, 1024, 0, st>>>(tile_sums, ntiles, out, n);
    k_scan_add<<<ntiles, SCAN_THREADS, 0, st>>>(out, tile_sums, n);
}

// ------------------------------------------------------------------ MSM kernels
template <class C>
static __global__ void __launch_bounds__(256) k_msm_digits(MsmPlan p, const uint32_t* __restrict__ scalars, const Affine<C>* __restrict__ points,
                                                    int32_t* __restrict__ digits, uint32_t* __restrict__ counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    constexpr int SL = C::SCALAR_LIMBS;
    const uint4* sp = reinterpret_cast<const uint4*>(scalars) + (size_t)i * (SL / 4);
    uint32_t sl[SL];
#pragma unroll
    for (int k = 0; k < SL / 4; k++) {
        uint4 a = __ldg(sp + k);
        sl[4 * k] = a.x;
        sl[4 * k + 1] = a.y;
        sl[4 * k + 2] = a.z;
        sl[4 * k + 3] = a.w;
    }
    typename C::F x = ldg_struct(&points[i].x);
    msm_digits_body<SL>(p, i, sl, x.is_zero(), digits, counts);
}

static __global__ void __launch_bounds__(256) k_msm_scatter(MsmPlan p, const int32_t* __restrict__ digits, const uint32_t* __restrict__ bstart,
                                                     uint32_t* __restrict__ cursor, uint32_t* __restrict__ entries) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    msm_scatter_body(p, i, digits, bstart, cursor, entries);
}

// Task-length histogram / placement go through a per-block shared-memory histogram so that the
// few hot lengths cost one global atomic per block instead of one per task.
constexpr uint32_t LEN_SH = 1026;   // shared histogram capacity (task caps above it use global atomics)

static __global__ void __launch_bounds__(256) k_msm_filltasks(MsmPlan p, const uint32_t* __restrict__ bstart, const uint32_t* __restrict__ tstart,
                                                              uint32_t* __restrict__ task_bucket, uint32_t* len_hist, uint32_t* __restrict__ heavy, uint32_t* heavy_count) {
    __shared__ uint32_t sh[LEN_SH];
    const bool use_sh = p.T + 2 <= LEN_SH;
    if (use_sh) {
        for (uint32_t l = threadIdx.x; l < p.T + 2; l += blockDim.x) sh[l] = 0;
        __syncthreads();
    }
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < p.nbuckets) msm_filltasks_body(p, b, bstart, tstart, task_bucket, use_sh ? sh : len_hist, heavy, heavy_count);
    if (use_sh) {
        __syncthreads();
        for (uint32_t l = threadIdx.x; l < p.T + 2; l += blockDim.x)
            if (sh[l]) atomicAdd(&len_hist[l], sh[l]);
    }
}

static __global__ void k_msm_task_offsets(MsmPlan p, const uint32_t* __restrict__ len_hist, uint32_t* __restrict__ len_cursor) {
    if (threadIdx.x == 0 && blockIdx.x == 0) msm_task_offsets_body(p, len_hist, len_cursor);
}

static __global__ void __launch_bounds__(256) k_msm_task_order(MsmPlan p, const uint32_t* __restrict__ bstart, const uint32_t* __restrict__ tstart,
                                                               const uint32_t* __restrict__ task_bucket, uint32_t* len_cursor, uint32_t* __restrict__ order) {
    __shared__ uint32_t sh[LEN_SH];
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ntasks = tstart[p.nbuckets];
    if (p.T + 2 > LEN_SH) {
        if (t < ntasks) msm_task_order_body(p, t, bstart, tstart, task_bucket, len_cursor, order);
        return;
    }
    for (uint32_t l = threadIdx.x; l < p.T + 2; l += blockDim.x) sh[l] = 0;
    __syncthreads();
    uint32_t len = 0, pos = 0;
    if (t < ntasks) {
        uint32_t b = task_bucket[t];
        uint32_t rem = bstart[b + 1] - bstart[b] - (t - tstart[b]) * p.T;
        len = rem < p.T ? rem : p.T;
        pos = atomicAdd(&sh[len], 1u);
    }
    __syncthreads();
    for (uint32_t l = threadIdx.x; l < p.T + 2; l += blockDim.x) {
        uint32_t c = sh[l];
        if (c) sh[l] = atomicAdd(&len_cursor[l], c);
    }
    __syncthreads();
    if (t < ntasks) order[sh[len] + pos] = t;
}

// The dominant kernel: one thread per task, points gathered with 128-bit loads, XYZZ
// accumulator in registers.
template <class C, int THREADS, int MINB>
static __global__ void __launch_bounds__(THREADS, MINB) k_msm_accumulate(MsmPlan p, const uint32_t* __restrict__ bstart, const uint32_t* __restrict__ tstart,
                                                                  const uint32_t* __restrict__ task_bucket, const uint32_t* __restrict__ order,
                                                                  const uint32_t* __restrict__ entries,
                                                                  const Affine<C>* __restrict__ points, XYZZ<C>* __restrict__ partials) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ntasks = tstart[p.nbuckets];
    if (i >= ntasks) return;
    uint32_t t = order[i];          // tasks sorted by length, longest first: lanes stay in step
    uint32_t b = task_bucket[t];
    uint32_t j = t - tstart[b];
    uint32_t lo = bstart[b] + j * p.T;
    uint32_t hi = bstart[b + 1];
    if (hi > lo + p.T) hi = lo + p.T;
    XYZZ<C> acc = XYZZ<C>::inf();
    for (uint32_t k = lo; k < hi; k++) {
        uint32_t e = __ldg(entries + k);
        Affine<C> q = ldg_struct(points + (e & 0x7FFFFFFFu));
        if (e >> 31) q.y = neg(q.y);
        if constexpr (sizeof(typename C::F) > 64) acc = pmadd_call<C>(acc, q);   // 24-limb fields: keep the kernel small
        else acc = pmadd<C>(acc, q);
    }
    stg_struct(partials + t, acc);
}

// Buckets split over several tasks.  2..16 partial sums: one thread per bucket adds them up
// serially (full lane efficiency; the common case when the task cap is below the bucket size).
// The sum lands in the bucket's first partial slot.
template <class C>
static __global__ void __launch_bounds__(128) k_msm_combine_small(MsmPlan p, const uint32_t* __restrict__ tstart, const uint32_t* __restrict__ heavy,
                                                                  const uint32_t* __restrict__ heavy_count, XYZZ<C>* partials) {
    uint32_t nh = heavy_count[0];
    for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < nh; h += gridDim.x * blockDim.x) {
        uint32_t b = heavy[h];
        uint32_t t0 = tstart[b], t1 = tstart[b + 1];
        XYZZ<C> acc = partials[t0];
        for (uint32_t t = t0 + 1; t < t1; t++) acc = padd_call<C>(acc, partials[t]);
        partials[t0] = acc;
    }
}

// More than 16 partial sums (a bucket that swallowed a large share of the input: all-equal scalars,
// the carry window of raw 256-bit scalars, the 1-valued signals of a witness): one 256-thread block
// per bucket, threads stride over the partials, shuffle tree per warp, then over the 8 warp sums --
// m/256 + 8 additions of depth instead of m.
template <class C>
static __global__ void __launch_bounds__(256) k_msm_combine_big(MsmPlan p, const uint32_t* __restrict__ tstart, const uint32_t* __restrict__ heavy,
                                                                const uint32_t* __restrict__ heavy_count, XYZZ<C>* partials) {
    __shared__ XYZZ<C> wsum[8];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t nh = heavy_count[1];
    for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
        uint32_t b = heavy[p.nbuckets + h];
        uint32_t t0 = tstart[b], t1 = tstart[b + 1];
        XYZZ<C> acc = XYZZ<C>::inf();
        for (uint32_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) acc = padd_call<C>(acc, partials[t]);
#pragma unroll 1
        for (int off = 16; off > 0; off >>= 1) {
            XYZZ<C> o = shfl_down_struct(acc, off);
            acc = padd_call<C>(acc, o);
        }
        if (lane == 0) wsum[wid] = acc;
        __syncthreads();
        if (wid == 0) {
            acc = lane < 8 ? wsum[lane] : XYZZ<C>::inf();
#pragma unroll 1
            for (int off = 4; off > 0; off >>= 1) {
                XYZZ<C> o = shfl_down_struct(acc, off);
                acc = padd_call<C>(acc, o);
            }
            if (lane == 0) partials[t0] = acc;
        }
        __syncthreads();
    }
}

// a level with at most this many (role, window, group) items runs the warp-cooperative form.  Measured at 2^20 / c = 20: with
// 4096 (level 3: 3072 warps) a proof is 0.5 ms shorter alone but 0.3 ms longer in a pipeline of proofs -- a cooperative addition
// holds the multiplier pipe 8x longer per useful product -- so only the levels that cannot fill a wave use it
constexpr uint32_t MSM_LEVEL_WARP_MAX = 1024;

template <class C>
static __global__ void __launch_bounds__(128, (sizeof(typename C::F) == 32 ? 4 : 1)) k_msm_segreduce(MsmPlan p, const uint32_t* __restrict__ tstart, const XYZZ<C>* __restrict__ partials,
                                                              XYZZ<C>* __restrict__ pool) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint32_t)p.nwin * p.S) return;
    msm_segreduce_body<C>(p, idx / p.S, idx % p.S, tstart, partials, pool);
}

// one level of the reduction tree: one thread per (role, window, group); roles are laid out role-major so that
// warps are role-uniform whenever nwin * Sl[l] is a multiple of 32
template <class C>
static __global__ void __launch_bounds__(128) k_msm_level(MsmPlan p, int l, XYZZ<C>* pool) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t per_role = (uint32_t)p.nwin * p.Sl[l];
    if (idx >= (uint32_t)l * per_role) return;
    int role = idx / per_role;
    uint32_t i = idx - role * per_role;
    msm_level_body<C>(p, l, i / p.Sl[l], i % p.Sl[l], role, pool);
}

// The same level with ONE WARP per (role, window, group) and warp-cooperative additions (warp_ec.cuh: the independent
// products of an addition on 4 lanes, 4 product levels deep instead of 14 sequential products).  Used for the small
// upper levels, where there are too few items to fill the machine and the level is pure dependent-chain latency: a
// cooperative addition has ~1/3 of the latency of a sequential one at 8x the pipe time, which is free up there.
template <class C>
static __global__ void __launch_bounds__(128) k_msm_level_warp(MsmPlan p, int l, XYZZ<C>* pool) {
    const uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // warp-uniform
    const uint32_t per_role = (uint32_t)p.nwin * p.Sl[l];
    if (idx >= (uint32_t)l * per_role) return;
    const int role = idx / per_role, lane = threadIdx.x & 31;
    const uint32_t i = idx - role * per_role, w = i / p.Sl[l], g = i % p.Sl[l];
    const uint32_t G = p.G[l], Sin = p.Sl[l - 1], Sout = p.Sl[l];
    const size_t in_per = (size_t)Sin * p.nwin, out_per = (size_t)Sout * p.nwin;
    const size_t ib = (size_t)w * Sin + (size_t)g * G, ob = (size_t)w * Sout + g;
    if (role > 0) {
        const XYZZ<C>* in = pool + p.offC[l - 1] + (size_t)(role - 1) * in_per;
        XYZZ<C> acc = in[ib];
#pragma unroll 1
        for (uint32_t j = 1; j < G; j++) acc = padd_warp4<C>(acc, in[ib + j]);
        if (lane == 0) pool[p.offC[l] + (size_t)(role - 1) * out_per + ob] = acc;
        return;
    }
    const XYZZ<C>* inT = pool + p.offT[l - 1];
    XYZZ<C> run = XYZZ<C>::inf(), acc = XYZZ<C>::inf();
#pragma unroll 1
    for (int j = (int)G - 1; j >= 1; j--) {
        run = padd_warp4<C>(run, inT[ib + j]);
        acc = padd_warp4<C>(acc, run);
    }
    run = padd_warp4<C>(run, inT[ib]);
    if (lane == 0) {
        pool[p.offC[l] + (size_t)(l - 1) * out_per + ob] = acc;
        pool[p.offT[l] + ob] = run;
    }
}

// window value from the plane sums (msm_window_value) with warp-cooperative doublings / additions; all lanes return it
template <class C>
__device__ __forceinline__ XYZZ<C> msm_window_value_warp(const MsmPlan& p, uint32_t w, const XYZZ<C>* bits) {
    const XYZZ<C>* bw = bits + (size_t)w * msm_nplanes(p);
    XYZZ<C> acc = XYZZ<C>::inf();
#pragma unroll 1
    for (int k = p.logSK - 1; k >= 0; k--) {
        acc = pdbl_warp4<C>(acc);
        acc = padd_warp4<C>(acc, bw[p.nlev + k]);
    }
#pragma unroll 1
    for (int l = p.nlev; l >= 1; l--) {
        for (uint32_t x = p.G[l]; x > 1; x >>= 1) acc = pdbl_warp4<C>(acc);
        acc = padd_warp4<C>(acc, bw[l - 1]);
    }
    return acc;
}

// plane sums over the top level (at most 32 groups per window): grid (nwin, nplanes), one warp each.
// plane r < nlev: sum of carried array r; plane nlev + k: sum of the T's of groups with bit k set.
template <class C>
static __global__ void __launch_bounds__(32) k_msm_bitsum(MsmPlan p, const XYZZ<C>* __restrict__ pool, XYZZ<C>* __restrict__ bits) {
    const uint32_t w = blockIdx.x;
    const int k = blockIdx.y, lane = threadIdx.x;
    const uint32_t SK = p.Sl[p.nlev];
    const size_t per = (size_t)SK * p.nwin;
    XYZZ<C> acc = XYZZ<C>::inf();
    for (uint32_t g = lane; g < SK; g += 32) {
        size_t o = (size_t)w * SK + g;
        if (k < p.nlev) acc = padd_call<C>(acc, pool[p.offC[p.nlev] + (size_t)k * per + o]);
        else if ((g >> (k - p.nlev)) & 1) acc = padd_call<C>(acc, pool[p.offT[p.nlev] + o]);
    }
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) {
        XYZZ<C> o = shfl_down_struct(acc, off);
        acc = padd_call<C>(acc, o);
    }
    if (lane == 0) bits[(size_t)w * msm_nplanes(p) + k] = acc;
}

// one block: thread w builds its window value, thread 0 runs the Horner over windows.
// out_xyzz: the sum (device consumers); out_jac: an un-normalised Jacobian triple X|Y|Z of the sum,
// Montgomery -- the reference's g1m_multiexp result format (callers normalise with g1_affine).
template <class C>
static __global__ void __launch_bounds__(160) k_msm_final(MsmPlan p, const XYZZ<C>* __restrict__ bits, XYZZ<C>* R, XYZZ<C>* out_xyzz, Jac<C>* out_jac) {
    if (p.nwin <= (int)(blockDim.x >> 5)) {       // few windows (one with a window table): one warp per window, cooperative chain
        const int w = threadIdx.x >> 5;
        if (w < p.nwin) {
            XYZZ<C> v = msm_window_value_warp<C>(p, w, bits);
            if ((threadIdx.x & 31) == 0) R[w] = v;
        }
    } else {
        for (int w = threadIdx.x; w < p.nwin; w += blockDim.x) R[w] = msm_window_value<C>(p, w, bits);
    }
    __syncthreads();
    if (threadIdx.x < 32) {      // one full warp: the doublings are 3-lane cooperative (warp_ec.cuh)
        MsmPlanLite pl{p.nwin, p.c};
        XYZZ<C> res = msm_horner_windows_warp<C>(pl, R);
        if (threadIdx.x == 0) {
            if (out_xyzz) *out_xyzz = res;
            if (out_jac) *out_jac = to_ext<C>(res);
        }
    }
}

// Sum of `count` points given as Jacobian triples; result an (un-normalised) Jacobian triple.
template <class C>
static __global__ void __launch_bounds__(32) k_sum_points(const Jac<C>* __restrict__ in, int count, Jac<C>* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    XYZZ<C> acc = XYZZ<C>::inf();
    for (int i = 0; i < count; i++) acc = padd_call<C>(acc, from_ext<C>(in[i]));
    *out = to_ext<C>(acc);
}

// g?m_affine (+ optional g?m_fromMontgomery) over an array of Jacobian triples, one thread each.
template <class C>
static __global__ void __launch_bounds__(64) k_affine_points(const Jac<C>* __restrict__ in, Jac<C>* __restrict__ out, int count, int from_montgomery) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Jac<C> a = to_affine<C>(from_ext<C>(in[i]));
    if (from_montgomery) {
        a.x = from_mont(a.x);
        a.y = from_mont(a.y);
        a.z = from_mont(a.z);
    }
    out[i] = a;
}

// ------------------------------------------------------------------ window tables (resident keys)
// table[w * n + i] = 2^(c w) * P_i as an affine point (x = 0: infinity), w < W.  One thread per point:
// the Jacobian chain P, 2^c P, 2^2c P, ... (dbl-2009-l), then ONE field inversion for all W normalisations
// (Montgomery's trick over the Z coordinates).  Built once when a proving key is loaded; with it every
// digit window of an MSM feeds the same bucket set and the result needs no doubling at all.
constexpr int MSM_TABLE_MAXW = 33;      // c >= 8 on 256-bit scalars

template <class C>
static __global__ void __launch_bounds__(64) k_msm_build_table(const Affine<C>* __restrict__ points, uint32_t n, int c, int W, Affine<C>* __restrict__ table) {
    static_assert(C::A_ZERO, "window tables are built with the a = 0 doubling");
    using F = typename C::F;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<C> p = ldg_struct(points + i);
    const Affine<C> zero{F::zero(), F::zero()};
    if (p.is_inf()) {
        for (int w = 0; w < W; w++) table[(size_t)w * n + i] = zero;
        return;
    }
    Jac<C> J[MSM_TABLE_MAXW];
    F pref[MSM_TABLE_MAXW];
    J[0] = Jac<C>{p.x, p.y, F::one()};
    pref[0] = F::one();
    int live = 1;                              // windows with a finite point
    for (int w = 1; w < W; w++) {
        Jac<C> t = J[w - 1];
#pragma unroll 1
        for (int k = 0; k < c; k++) t = jdbl_a0<C>(t);
        if (t.z.is_zero()) break;              // a point of 2-power order: everything above is infinity
        J[w] = t;
        pref[w] = mul(pref[w - 1], t.z);
        live = w + 1;
    }
    F inv_all = inv(pref[live - 1]);
    for (int w = live - 1; w >= 1; w--) {
        F zi = mul(inv_all, pref[w - 1]);      // 1 / Z_w
        inv_all = mul(inv_all, J[w].z);
        F zi2 = sqr(zi);
        table[(size_t)w * n + i] = Affine<C>{mul(J[w].x, zi2), mul(J[w].y, mul(zi2, zi))};
    }
    table[i] = p;
    for (int w = live; w < W; w++) table[(size_t)w * n + i] = zero;
}

// ------------------------------------------------------------------ workspace + launch
struct MsmWorkspace {
    int32_t* digits;
    uint32_t *counts, *cursor, *bstart, *tstart, *tile_sums, *entries, *task_bucket, *order, *heavy, *small_ctrs;   // small_ctrs: [0..1] heavy counts, [8..] len_hist, [8+T+2..] len_cursor
    void *partials, *pool, *bits, *R;
};

// Carves the workspace out of one device buffer; returns bytes needed (base may be null to size).
template <class C>
inline size_t msm_carve(const MsmPlan& p, uint8_t* base, MsmWorkspace* ws) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        uint8_t* ptr = base ? base + off : nullptr;
        off += align256(bytes);
        return ptr;
    };
    size_t entries = (size_t)p.ndig * p.n;
    uint8_t* d = take(entries * sizeof(int32_t));
    uint8_t* cn = take((size_t)p.nbuckets * 4);
    uint8_t* cu = take((size_t)p.nbuckets * 4);
    uint8_t* bs = take(((size_t)p.nbuckets + 1) * 4);
    uint8_t* ts = take(((size_t)p.nbuckets + 1) * 4);
    uint8_t* tl = take(((size_t)p.nbuckets / SCAN_TILE + 2) * 4);
    uint8_t* en = take((entries + 1) * 4);
    uint8_t* tb = take((size_t)p.max_tasks * 4);
    uint8_t* od = take((size_t)p.max_tasks * 4);
    uint8_t* hl = take((size_t)p.nbuckets * 8);
    uint8_t* hc = take((8 + 2 * ((size_t)p.T + 2)) * 4);
    uint8_t* pa = take((size_t)p.max_tasks * sizeof(XYZZ<C>));
    uint8_t* po = take(((size_t)p.pool_elems + 1) * sizeof(XYZZ<C>));
    uint8_t* bi = take((size_t)p.nwin * msm_nplanes(p) * sizeof(XYZZ<C>));
    uint8_t* rr = take((size_t)p.nwin * sizeof(XYZZ<C>));
    if (ws) {
        ws->digits = (int32_t*)d;
        ws->counts = (uint32_t*)cn;
        ws->cursor = (uint32_t*)cu;
        ws->bstart = (uint32_t*)bs;
        ws->tstart = (uint32_t*)ts;
        ws->tile_sums = (uint32_t*)tl;
        ws->entries = (uint32_t*)en;
        ws->task_bucket = (uint32_t*)tb;
        ws->order = (uint32_t*)od;
        ws->heavy = (uint32_t*)hl;
        ws->small_ctrs = (uint32_t*)hc;
        ws->partials = pa;
        ws->pool = po;
        ws->bits = bi;
        ws->R = rr;
    }
    return off;
}

// Window size / task cap for n pairs on a 148-SM part.
inline void msm_choose(uint32_t n, int* c, uint32_t* T, int sbits = 256, int table = 0) {
    int lg = 0;
    while ((1ull << (lg + 1)) <= n) lg++;
    int cc = lg - 4;
    if (cc < 4) cc = 4;
    if (cc > 16) cc = 16;
    if (table) {      // one shared bucket set: the reduce cost is paid once, so wider windows (fewer additions per point) win
        cc = lg;
        if (cc < 8) cc = 8;
        if (cc > 20) cc = 20;
    }
    *c = cc;
    int nwin = (sbits + 1 + cc - 1) / cc;
    uint64_t entries = (uint64_t)nwin * n;
    uint64_t want_tasks = 148ull * 512ull * 4ull;     // a few waves of resident threads
    uint64_t t = entries / want_tasks;
    if (t < 8) t = 8;
    if (t > 64) t = 64;
    *T = (uint32_t)t;
}

template <class C>
inline cudaError_t msm_launch(const MsmPlan& p, const MsmWorkspace& w, const uint32_t* d_scalars, const Affine<C>* d_points,
                              XYZZ<C>* d_out_xyzz, Jac<C>* d_out_jac, cudaStream_t st, StageTimer* tm = nullptr, unsigned long long* nlaunch = nullptr, int acc_variant = 0) {
    using X = XYZZ<C>;
    auto mark = [&](const char* nm) { if (tm) tm->mark(nm, st); };
    cudaMemsetAsync(w.counts, 0, (size_t)p.nbuckets * 4, st);
    uint32_t* heavy_count = w.small_ctrs;
    uint32_t* len_hist = w.small_ctrs + 8;
    uint32_t* len_cursor = len_hist + p.T + 2;
    cudaMemsetAsync(w.small_ctrs, 0, (8 + 2 * ((size_t)p.T + 2)) * 4, st);
    if (p.n) k_msm_digits<C><<<(p.n + 255) / 256, 256, 0, st>>>(p, d_scalars, d_points, w.digits, w.counts);
    mark("digits");
    launch_scan(w.counts, w.bstart, w.tile_sums, p.nbuckets, 0, 1, st);
    launch_scan(w.counts, w.tstart, w.tile_sums, p.nbuckets, 1, p.T, st);
    cudaMemcpyAsync(w.cursor, w.counts, (size_t)p.nbuckets * 4, cudaMemcpyDeviceToDevice, st);
    mark("scan");
    if (p.n) k_msm_scatter<<<(p.n + 255) / 256, 256, 0, st>>>(p, w.digits, w.bstart, w.cursor, w.entries);
    mark("scatter");
    k_msm_filltasks<<<(p.nbuckets + 255) / 256, 256, 0, st>>>(p, w.bstart, w.tstart, w.task_bucket, len_hist, w.heavy, heavy_count);
    k_msm_task_offsets<<<1, 32, 0, st>>>(p, len_hist, len_cursor);
    k_msm_task_order<<<(p.max_tasks + 255) / 256, 256, 0, st>>>(p, w.bstart, w.tstart, w.task_bucket, len_cursor, w.order);
    mark("tasks");
    constexpr int AT = sizeof(typename C::F) > 32 ? 128 : 256;
    constexpr int AB = sizeof(typename C::F) > 32 ? 1 : 2;
    auto acc_args = [&](auto kern, int threads) {
        kern<<<(p.max_tasks + threads - 1) / threads, threads, 0, st>>>(p, w.bstart, w.tstart, w.task_bucket, w.order, w.entries, d_points, (X*)w.partials);
    };
    bool launched = false;
    // acc_variant: low 4 bits select the variant of the 8-limb base-field curves, bits 4..7 the one of the Fq2 curves
    if constexpr (sizeof(typename C::F) == 64) acc_variant = (acc_variant >> 4) & 15;
    else acc_variant &= 15;
    if constexpr (sizeof(typename C::F) == 32) {
        if (acc_variant == 1) { acc_args(k_msm_accumulate<C, 128, 5>, 128); launched = true; }        // ~100 regs, 20 warps/SM
        else if (acc_variant == 2) { acc_args(k_msm_accumulate<C, 128, 4>, 128); launched = true; }   // 128 regs, smaller blocks
        else if (acc_variant == 3) { acc_args(k_msm_accumulate<C, 192, 3>, 192); launched = true; }   // 113 regs, 18 warps/SM
    } else if constexpr (sizeof(typename C::F) == 64) {
        if (acc_variant == 1) { acc_args(k_msm_accumulate<C, 64, 3>, 64); launched = true; }          // G2: 168 regs x 3 blocks of 2 warps
        else if (acc_variant == 2 || acc_variant == 3) { acc_args(k_msm_accumulate<C, 128, 3>, 128); launched = true; }   // G2: 168 regs, 12 warps/SM, spills
    }
    if (!launched) acc_args(k_msm_accumulate<C, AT, AB>, AT);
    mark("accumulate");
    k_msm_combine_small<C><<<148 * 8, 128, 0, st>>>(p, w.tstart, w.heavy, heavy_count, (X*)w.partials);
    k_msm_combine_big<C><<<148, 256, 0, st>>>(p, w.tstart, w.heavy, heavy_count, (X*)w.partials);
    mark("combine");
    uint32_t nseg = (uint32_t)p.nwin * p.S;
    k_msm_segreduce<C><<<(nseg + 127) / 128, 128, 0, st>>>(p, w.tstart, (const X*)w.partials, (X*)w.pool);
    mark("segreduce");
    for (int l = 2; l <= p.nlev; l++) {
        uint32_t nt = (uint32_t)l * (uint32_t)p.nwin * p.Sl[l];
        if (nt <= MSM_LEVEL_WARP_MAX) k_msm_level_warp<C><<<(nt + 3) / 4, 128, 0, st>>>(p, l, (X*)w.pool);     // one warp per item
        else k_msm_level<C><<<(nt + 127) / 128, 128, 0, st>>>(p, l, (X*)w.pool);
    }
    k_msm_bitsum<C><<<dim3(p.nwin, msm_nplanes(p)), 32, 0, st>>>(p, (const X*)w.pool, (X*)w.bits);
    mark("bitsum");
    k_msm_final<C><<<1, 160, 0, st>>>(p, (const X*)w.bits, (X*)w.R, d_out_xyzz, d_out_jac);
    mark("final");
    if (nlaunch) *nlaunch += (p.n ? 2 : 0) + 6 + 3 + 1 + 2 + 1 + (p.nlev - 1) + 1 + 1;   // digits+scatter, 2 scans, task table, accumulate, combine x2, segreduce, levels, bitsum, final
    return cudaGetLastError();
}

}  // namespace ws
