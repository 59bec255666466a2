// prover.cu -- NTT entry points, the H-polynomial pipeline (worker command CALC_H) and the
// Groth16 prover (Bn128.groth16GenProof) on the GPU, behind the C ABI.
//
// Reference call stack replaced here (paths under /root/reference):
//   Bn128.groth16GenProof(signals, pkey)            src/bn128.js:580-720
//     header + section slicing                      src/bn128.js:581-604   (layout: tools/buildpkey.js:130-183)
//     calcH -> worker CALC_H                        src/bn128.js:569-578, :126-166
//        fft_toMontgomeryN                          src/build_fft.js:418-458
//        pol_zero, pol_constructLC                  src/build_pol.js:26-143
//        fft_copyNInterleaved, fft_mulN             src/build_fft.js:374-414, :461-505
//        fft_ifft / fft_fft(odd)                    src/build_fft.js:159-221
//        fft_fromMontgomeryN                        src/build_fft.js:507-547
//     g1_multiexp x4, g2_multiexp x1                src/bn128.js:614-620
//     final assembly with r, s                      src/bn128.js:642-712
#include <cuda_runtime.h>
#include <sys/random.h>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/wasmsnark_b200.h"
#include "ctx.cuh"
#include "ec.cuh"
#include "msm_api.cuh"
#include "ntt.cuh"
#include "warp_ec.cuh"

#define WS_API __attribute__((visibility("default")))

using namespace ws;

// ============================================================================ element-wise kernels
static __global__ void __launch_bounds__(128) k_fr_to_mont(const Fr* __restrict__ in, Fr* __restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fr_st(out + i, to_mont(fr_ld(in + i)));
}
static __global__ void __launch_bounds__(128) k_fr_from_mont(const Fr* __restrict__ in, Fr* __restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fr_st(out + i, from_mont(fr_ld(in + i)));
}

// pol_constructLC in row form: out[row] = sum over the row's terms of signalM[sig] * coef.
// The key stores the terms per signal (src/build_pol.js:61-143 walks them signal-major and
// accumulates into res[idx]); they are regrouped per row once at key load, which turns the
// scatter-add into a conflict-free gather.  Field addition is commutative, so the sum is the same.
static __global__ void __launch_bounds__(128) k_csr_lc(const uint32_t* __restrict__ row_start, const uint32_t* __restrict__ sig,
                                                       const Fr* __restrict__ coef, const Fr* __restrict__ signalsM, Fr* __restrict__ out, uint32_t nrows) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    Fr acc = Fr::zero();
    uint32_t lo = row_start[r], hi = row_start[r + 1];
    for (uint32_t k = lo; k < hi; k++) acc = add(acc, mul(fr_ldg(signalsM + sig[k]), fr_ldg(coef + k)));
    fr_st(out + r, acc);
}

// P[2i] = a_even[i] * b_even[i]; P[2i+1] = a_odd[i] * b_odd[i]
// (fft_copyNInterleaved x4 + fft_mulN over the 2n domain, src/bn128.js:146-157)
static __global__ void __launch_bounds__(128) k_interleave_mul(const Fr* __restrict__ ae, const Fr* __restrict__ be, const Fr* __restrict__ ao,
                                                               const Fr* __restrict__ bo, Fr* __restrict__ P, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_st(P + 2 * (size_t)i, mul(fr_ld(ae + i), fr_ld(be + i)));
    fr_st(P + 2 * (size_t)i + 1, mul(fr_ld(ao + i), fr_ld(bo + i)));
}

// ============================================================================ final assembly
struct FixedPoints {      // alfa1, beta1, delta1 (64 B each), beta2, delta2 (128 B each): pkey bytes 40..488
    G1Affine alfa1, beta1, delta1;
    G2Affine beta2, delta2;
};

struct AssembleArgs {
    const G1X *A, *B1, *C, *H;
    const G2X* B2;
    const FixedPoints* fx;
    const uint8_t* rs;          // r (32 bytes) then s (32 bytes), raw LE
    Jac<G1Curve>* pi_a;         // outputs: affine, out of Montgomery form (src/bn128.js:706-712)
    Jac<G2Curve>* pi_b;
    Jac<G1Curve>* pi_c;
};

template <class C>
__device__ __forceinline__ XYZZ<C> load_fixed(const Affine<C>& p) {   // loadPoint1/2: Z = 1 always (src/bn128.js:441-453)
    using F = typename C::F;
    return XYZZ<C>{p.x, p.y, F::one(), F::one()};
}

template <class C>
__device__ __forceinline__ Jac<C> affine_std(const XYZZ<C>& p) {   // g?m_affine + g?m_fromMontgomery
    Jac<C> a = to_affine<C>(p);
    a.x = from_mont(a.x);
    a.y = from_mont(a.y);
    a.z = from_mont(a.z);
    return a;
}

// The four products that depend only on (r, s) and the key's fixed points: r*delta1, s*delta1,
// (r*s)*delta1 (512-bit scalar) and s*delta2.  One block, 4 warps, lane 0 of each warp runs one
// product; launched on a side stream so it overlaps the multi-scalar multiplications.
struct PreProducts {
    G1X rd1, sd1, rsd1;
    G2X sd2;
};

// Fixed-base comb tables of the key's delta1 / delta2: entry [w][k] = k * 2^(4w) * delta (affine), k = 1..15.
// With them a product k*delta is a sum of one table entry per nibble of k -- no doublings -- and the nibbles
// are spread over the lanes of a warp, so the (r, s)-only products cost ~10 additions of depth.
constexpr int D1_WINDOWS = 128;     // 512-bit scalars (r*s)
constexpr int D2_WINDOWS = 64;      // 256-bit scalars
struct FixedBaseTables {
    G1Affine d1[D1_WINDOWS][16];
    G2Affine d2[D2_WINDOWS][16];
};

template <class C>
__device__ __forceinline__ void comb_rows(const Affine<C>& base_pt, int w, Affine<C>* row) {
    using F = typename C::F;
    Jac<C> j{base_pt.x, base_pt.y, F::one()};
    for (int k = 0; k < 4 * w; k++) j = jdbl_a0<C>(j);          // 2^(4w) * delta
    XYZZ<C> b = from_jac<C>(j), acc = b;
    row[0] = Affine<C>{F::zero(), F::zero()};
    for (int k = 1; k < 16; k++) {
        Jac<C> a = to_affine<C>(acc);
        row[k] = acc.is_inf() ? Affine<C>{F::zero(), F::zero()} : Affine<C>{a.x, a.y};
        acc = padd_call<C>(acc, b);
    }
}

static __global__ void __launch_bounds__(64) k_build_fixed_tables(const FixedPoints* __restrict__ fx, FixedBaseTables* t) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < D1_WINDOWS) comb_rows<G1Curve>(fx->delta1, w, t->d1[w]);
    else if (w < D1_WINDOWS + D2_WINDOWS) comb_rows<G2Curve>(fx->delta2, w - D1_WINDOWS, t->d2[w - D1_WINDOWS]);
}

// sum over the nibbles of k (len bytes) of row[w][nibble]: lanes stride over the windows, shuffle tree.
template <class C, int WINDOWS>
__device__ __forceinline__ XYZZ<C> comb_product_warp(const Affine<C> (*rows)[16], const uint8_t* k, int len) {
    const int lane = threadIdx.x & 31;
    XYZZ<C> acc = XYZZ<C>::inf();
    for (int w = lane; w < 2 * len && w < WINDOWS; w += 32) {
        uint32_t nib = (k[w >> 1] >> ((w & 1) * 4)) & 15u;
        if (nib) {
            Affine<C> q = rows[w][nib];
            if (!q.is_inf()) acc = pmadd_call<C>(acc, q);
        }
    }
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) {
        XYZZ<C> o = shfl_bcast_struct(acc, (lane + off) & 31);     // every lane keeps a full copy of its subtree
        if (lane < off) acc = padd_call<C>(acc, o);
    }
    return acc;       // valid on lane 0
}

static __global__ void __launch_bounds__(128) k_assemble_pre(const FixedPoints* __restrict__ fx, const FixedBaseTables* __restrict__ tb,
                                                            const uint8_t* __restrict__ rs, PreProducts* out) {
    __shared__ uint8_t s_rs[64], s_prod[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 64) s_rs[threadIdx.x] = rs[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        // int_mul(r, s): exact 512-bit product (src/bn128.js:701, src/build_int.js:285-380)
        uint32_t r32[8], s32[8], p[16];
        memcpy(r32, s_rs, 32);
        memcpy(s32, s_rs + 32, 32);
        for (int i = 0; i < 16; i++) p[i] = 0;
        for (int i = 0; i < 8; i++) {
            uint64_t c = 0;
            for (int j = 0; j < 8; j++) {
                c += (uint64_t)r32[i] * s32[j] + p[i + j];
                p[i + j] = (uint32_t)c;
                c >>= 32;
            }
            p[i + 8] = (uint32_t)c;
        }
        memcpy(s_prod, p, 64);
    }
    __syncthreads();
    const uint8_t* r = s_rs;
    const uint8_t* s = s_rs + 32;
    if (tb) {         // resident key: fixed-base comb, no doublings
        if (warp < 3) {
            G1X v = comb_product_warp<G1Curve, D1_WINDOWS>(tb->d1, warp == 0 ? r : (warp == 1 ? s : s_prod), warp == 2 ? 64 : 32);
            if (lane == 0) (warp == 0 ? out->rd1 : (warp == 1 ? out->sd1 : out->rsd1)) = v;
        } else {
            G2X v = comb_product_warp<G2Curve, D2_WINDOWS>(tb->d2, s, 32);
            if (lane == 0) out->sd2 = v;
        }
        return;
    }
    // every lane of a warp carries the same state (warp-cooperative doublings); lane 0 stores
    if (warp < 3) {
        G1X v = times_scalar_w4_warp<G1Curve>(load_fixed<G1Curve>(fx->delta1), warp == 0 ? r : (warp == 1 ? s : s_prod), warp == 2 ? 64 : 32);
        if (lane == 0) (warp == 0 ? out->rd1 : (warp == 1 ? out->sd1 : out->rsd1)) = v;
    } else {
        G2X v = times_scalar_w4_warp<G2Curve>(load_fixed<G2Curve>(fx->delta2), s, 32);
        if (lane == 0) out->sd2 = v;
    }
}

// After the multi-scalar multiplications: 4 warps, three phases.
//   phase 1  warp 0: pi_a = A + alfa1 + r*delta1          (src/bn128.js:670-673)
//            warp 1: pib1 = B1 + beta1 + s*delta1         (:680-683)
//            warp 2: pi_b = B2 + beta2 + s*delta2         (:675-678)
//            warp 3: C + H - (r*s)*delta1                 (:686-688, :699-704)
//   phase 2  warp 0: s*pi_a   warp 1: r*pib1              (:691, :695; warp-cooperative scalar multiplications)
//            warp 2: pi_b -> affine, fromMontgomery       warp 3: pi_a -> affine  (:706-712; inversions off the critical path)
//   phase 3  pi_c = phase-1 sum + s*pi_a + r*pib1, affine + fromMontgomery
static __global__ void __launch_bounds__(128) k_assemble_post(AssembleArgs g, const PreProducts* __restrict__ pre) {
    __shared__ G1X s_pia, s_pib1, s_spia, s_rpib1, s_pic0;
    __shared__ G2X s_pib;
    __shared__ uint8_t s_rs[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 64) s_rs[threadIdx.x] = g.rs[threadIdx.x];
    const uint8_t* r = s_rs;
    const uint8_t* s = s_rs + 32;
    if (lane == 0) {
        if (warp == 0) s_pia = padd_call<G1Curve>(pre->rd1, padd_call<G1Curve>(load_fixed<G1Curve>(g.fx->alfa1), *g.A));
        else if (warp == 1) s_pib1 = padd_call<G1Curve>(pre->sd1, padd_call<G1Curve>(load_fixed<G1Curve>(g.fx->beta1), *g.B1));
        else if (warp == 2) s_pib = padd_call<G2Curve>(pre->sd2, padd_call<G2Curve>(load_fixed<G2Curve>(g.fx->beta2), *g.B2));
        else s_pic0 = padd_call<G1Curve>(pneg<G1Curve>(pre->rsd1), padd_call<G1Curve>(*g.H, *g.C));
    }
    __syncthreads();
    if (warp == 0) {
        G1X v = times_scalar_w4_warp<G1Curve>(s_pia, s, 32);
        if (lane == 0) s_spia = v;
    } else if (warp == 1) {
        G1X v = times_scalar_w4_warp<G1Curve>(s_pib1, r, 32);
        if (lane == 0) s_rpib1 = v;
    } else if (lane == 0) {
        if (warp == 2) *g.pi_b = affine_std<G2Curve>(s_pib);
        else *g.pi_a = affine_std<G1Curve>(s_pia);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        G1X pic = padd_call<G1Curve>(s_spia, s_pic0);
        pic = padd_call<G1Curve>(s_rpib1, pic);
        *g.pi_c = affine_std<G1Curve>(pic);
    }
}

// ============================================================================ proving key
// Row-grouped form of one polynomial section (u32 ncoefs; {u32 idx; 32-byte coef} x ncoefs per signal;
// tools/buildpkey.js:79-89).  Returns false if the section runs past `len` or an index is out of range.
static bool build_csr(const uint8_t* sec, size_t len, uint32_t n_signals, uint32_t domain, std::vector<uint32_t>& row_start,
                      std::vector<uint32_t>& sig, std::vector<uint8_t>& coef) {
    row_start.assign((size_t)domain + 1, 0);
    size_t pos = 0, total = 0;
    for (uint32_t i = 0; i < n_signals; i++) {
        if (pos + 4 > len) return false;
        uint32_t nc;
        memcpy(&nc, sec + pos, 4);
        pos += 4;
        if ((size_t)nc * 36 > len - pos) return false;
        for (uint32_t k = 0; k < nc; k++) {
            uint32_t idx;
            memcpy(&idx, sec + pos, 4);
            if (idx >= domain) return false;
            row_start[idx + 1]++;
            pos += 36;
        }
        total += nc;
    }
    for (uint32_t r = 0; r < domain; r++) row_start[r + 1] += row_start[r];
    sig.resize(total + 1);
    coef.resize((total + 1) * 32);
    std::vector<uint32_t> cur(row_start.begin(), row_start.end() - 1);
    pos = 0;
    for (uint32_t i = 0; i < n_signals; i++) {
        uint32_t nc;
        memcpy(&nc, sec + pos, 4);
        pos += 4;
        for (uint32_t k = 0; k < nc; k++) {
            uint32_t idx;
            memcpy(&idx, sec + pos, 4);
            uint32_t at = cur[idx]++;
            sig[at] = i;
            memcpy(&coef[(size_t)at * 32], sec + pos + 4, 32);
            pos += 36;
        }
    }
    return true;
}

struct DevCsr {
    uint32_t* row_start = nullptr;
    uint32_t* sig = nullptr;
    Fr* coef = nullptr;
    void free_all() {
        cudaFree(row_start);
        cudaFree(sig);
        cudaFree(coef);
        row_start = sig = nullptr;
        coef = nullptr;
    }
};

static bool upload_csr(const std::vector<uint32_t>& rs, const std::vector<uint32_t>& sg, const std::vector<uint8_t>& cf, DevCsr& d, cudaStream_t st) {
    if (cudaMalloc(&d.row_start, rs.size() * 4) != cudaSuccess) return false;
    if (cudaMalloc(&d.sig, sg.size() * 4) != cudaSuccess) return false;
    if (cudaMalloc(&d.coef, cf.size()) != cudaSuccess) return false;
    cudaMemcpyAsync(d.row_start, rs.data(), rs.size() * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.sig, sg.data(), sg.size() * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.coef, cf.data(), cf.size(), cudaMemcpyHostToDevice, st);
    return cudaStreamSynchronize(st) == cudaSuccess;
}

struct ws_pkey {
    ws_ctx* ctx = nullptr;
    uint32_t n_vars = 0, n_public = 0, domain = 0;
    int logn = 0;
    // point-range shard held by this key object (rank r of w): contiguous ranges, floor(n/w) each and the
    // last rank takes the remainder -- the reference's worker split (src/bn128.js:353-361)
    int rank = 0, world = 1;
    // element range of each point section held by this key object, in section order A, B1, B2, C, H (SEC_*): A / B1 /
    // B2 index the signals, C the signals from nPublic+1 on (lo is relative to that), H the domain.  The uniform shard
    // gives A, B1, B2 the same range; ws_bn128_pkey_load_ranges sets each independently (balanced split, parallel.py).
    uint32_t lo[5] = {}, cnt[5] = {};
    DevCsr polA, polB;
    uint8_t* points = nullptr;     // one allocation: A | B1 | B2 | C | H, each 256-byte aligned
    G1Affine *A = nullptr, *B1 = nullptr, *C = nullptr, *H = nullptr;
    G2Affine* B2 = nullptr;
    FixedPoints* fixed = nullptr;
    FixedBaseTables* fixed_tables = nullptr;    // comb tables of delta1 / delta2 (built with the window tables)
    // window tables (2^(c w) * P_i, see k_msm_build_table): one per section, built at load; tc_* = window bits, 0 = none
    void* tab[5] = {};
    int tc[5] = {};
    const void* msm_points(int sec) const {
        if (tab[sec]) return tab[sec];
        return sec == 0 ? (const void*)A : sec == 1 ? (const void*)B1 : sec == 2 ? (const void*)B2 : sec == 3 ? (const void*)C : (const void*)H;
    }
};
enum { SEC_A = 0, SEC_B1 = 1, SEC_B2 = 2, SEC_C = 3, SEC_H = 4 };

static int log2_exact(uint32_t n) {
    if (n == 0 || (n & (n - 1))) return -1;
    int l = 0;
    while ((1u << l) < n) l++;
    return l;
}

static void pkey_release(ws_pkey* k) {
    if (!k) return;
    k->polA.free_all();
    k->polB.free_all();
    cudaFree(k->points);
    cudaFree(k->fixed);
    cudaFree(k->fixed_tables);
    for (int i = 0; i < 5; i++) cudaFree(k->tab[i]);
    delete k;
}

static void shard_range(uint32_t n, int rank, int world, uint32_t* lo, uint32_t* cnt) {
    uint32_t per = n / (uint32_t)world;
    *lo = per * (uint32_t)rank;
    *cnt = rank < world - 1 ? per : n - per * (uint32_t)(world - 1);
}

// Window tables for the five point sections (HBM is plentiful: 13 x the section size at 2^20).  Failure to
// allocate is not an error: the section then runs the table-free MSM.
static void build_key_tables(ws_ctx* c, ws_pkey* k) {
    const char* force = getenv("WS_TABLE_C");          // tuning hook: window bits of the tables (default: from the section size)
    const void* pts[5] = {k->A, k->B1, k->B2, k->C, k->H};
    for (int sec = 0; sec < 5; sec++) {
        k->tab[sec] = nullptr;
        k->tc[sec] = 0;
        const size_t cnt = k->cnt[sec], psize = sec == SEC_B2 ? 128 : 64;
        if (cnt == 0) continue;
        int cc = msm_table_choose_c(cnt);
        if (force && atoi(force) >= 8 && atoi(force) <= 22) cc = atoi(force);
        int W = msm_table_windows_n(cc);
        if ((uint64_t)W * cnt >= 0x80000000ull) continue;
        void* t = nullptr;
        if (cudaMalloc(&t, (size_t)W * cnt * psize) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        int rc = sec == SEC_B2 ? msm_g2_build_table(c, pts[sec], cnt, cc, t, c->stream) : msm_g1_build_table(c, pts[sec], cnt, cc, t, c->stream);
        if (rc || cudaStreamSynchronize(c->stream) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(t);
            continue;
        }
        k->tc[sec] = cc;
        k->tab[sec] = t;
    }
}

// ranges: nullptr = the uniform shard of (rank, world); else {lo, cnt} x 5 in section order A, B1, B2, C, H
static int pkey_load(ws_ctx* c, const uint8_t* pkey, size_t len, ws_pkey** out, int rank = 0, int world = 1, const uint32_t* ranges = nullptr) {
    if (world < 1 || rank < 0 || rank >= world) return c->fail(WS_E_ARG, "bad rank / world");
    if (len < 488) return c->fail(WS_E_ARG, "proving key shorter than its header");
    uint32_t h[10];
    memcpy(h, pkey, 40);
    const uint32_t nv = h[0], npub = h[1], dom = h[2];
    const size_t pA = h[3], pB = h[4], pPA = h[5], pPB1 = h[6], pPB2 = h[7], pPC = h[8], pPH = h[9];
    int logn = log2_exact(dom);
    if (logn < 0 || logn > 27) return c->fail(WS_E_SIZE, "domainSize must be a power of two <= 2^27");   // 2n-point transform <= 2^28
    if (nv == 0 || npub + 1 > nv) return c->fail(WS_E_ARG, "bad nVars / nPublic");
    const size_t nC = (size_t)nv - npub - 1;
    if (!(488 <= pA && pA <= pB && pB <= pPA && pPA + (size_t)nv * 64 <= len && pPB1 + (size_t)nv * 64 <= len && pPB2 + (size_t)nv * 128 <= len &&
          pPC + nC * 64 <= len && pPH + (size_t)dom * 64 <= len))
        return c->fail(WS_E_ARG, "proving key sections out of bounds");
    cudaSetDevice(c->device);
    const bool trace = getenv("WS_TRACE_LOAD") != nullptr;     // stage times of the key load on stderr
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        cudaStreamSynchronize(c->stream);
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[ws pkey_load] %-14s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    std::vector<uint32_t> rs, sg;
    std::vector<uint8_t> cf;
    ws_pkey* k = new ws_pkey();
    k->ctx = c;
    k->n_vars = nv;
    k->n_public = npub;
    k->domain = dom;
    k->logn = logn;
    k->rank = rank;
    k->world = world;
    const uint32_t sec_n[5] = {nv, nv, nv, (uint32_t)nC, dom};
    if (ranges) {
        for (int i = 0; i < 5; i++) {
            k->lo[i] = ranges[2 * i];
            k->cnt[i] = ranges[2 * i + 1];
            if ((uint64_t)k->lo[i] + k->cnt[i] > sec_n[i]) {
                pkey_release(k);
                return c->fail(WS_E_ARG, "section range out of bounds");
            }
        }
    } else {
        shard_range(nv, rank, world, &k->lo[SEC_A], &k->cnt[SEC_A]);
        k->lo[SEC_B1] = k->lo[SEC_B2] = k->lo[SEC_A];
        k->cnt[SEC_B1] = k->cnt[SEC_B2] = k->cnt[SEC_A];
        shard_range((uint32_t)nC, rank, world, &k->lo[SEC_C], &k->cnt[SEC_C]);
        shard_range(dom, rank, world, &k->lo[SEC_H], &k->cnt[SEC_H]);
    }
    // the reference slices polsA as [pPolsA, pPolsA + pPolsB) (src/bn128.js:592); the section proper ends at pPolsB.
    // The two sections are regrouped by row on two host threads.
    std::vector<uint32_t> rs2, sg2;
    std::vector<uint8_t> cf2;
    bool okA = false, okB = false;
    std::thread tb([&] { okB = build_csr(pkey + pB, pPA - pB, nv, dom, rs2, sg2, cf2); });
    okA = build_csr(pkey + pA, pB - pA, nv, dom, rs, sg, cf);
    tb.join();
    if (!okA || !upload_csr(rs, sg, cf, k->polA, c->stream)) {
        pkey_release(k);
        return c->fail(WS_E_ARG, "polsA section malformed (or out of device memory)");
    }
    if (!okB || !upload_csr(rs2, sg2, cf2, k->polB, c->stream)) {
        pkey_release(k);
        return c->fail(WS_E_ARG, "polsB section malformed (or out of device memory)");
    }
    lap("csr A+B");
    const size_t nA = k->cnt[SEC_A], nB1 = k->cnt[SEC_B1], nB2 = k->cnt[SEC_B2], cc = k->cnt[SEC_C], hc = k->cnt[SEC_H];
    size_t oA = 0, oB1 = oA + align256(nA * 64 + 64), oB2 = oB1 + align256(nB1 * 64 + 64), oC = oB2 + align256(nB2 * 128 + 64),
           oH = oC + align256(cc * 64 + 64), tot = oH + align256(hc * 64 + 64);
    if (cudaMalloc(&k->points, tot) != cudaSuccess || cudaMalloc(&k->fixed, sizeof(FixedPoints)) != cudaSuccess) {
        pkey_release(k);
        return c->fail_cuda("proving key upload");
    }
    k->A = (G1Affine*)(k->points + oA);
    k->B1 = (G1Affine*)(k->points + oB1);
    k->B2 = (G2Affine*)(k->points + oB2);
    k->C = (G1Affine*)(k->points + oC);
    k->H = (G1Affine*)(k->points + oH);
    if (nA) cudaMemcpyAsync(k->A, pkey + pPA + (size_t)k->lo[SEC_A] * 64, nA * 64, cudaMemcpyHostToDevice, c->stream);
    if (nB1) cudaMemcpyAsync(k->B1, pkey + pPB1 + (size_t)k->lo[SEC_B1] * 64, nB1 * 64, cudaMemcpyHostToDevice, c->stream);
    if (nB2) cudaMemcpyAsync(k->B2, pkey + pPB2 + (size_t)k->lo[SEC_B2] * 128, nB2 * 128, cudaMemcpyHostToDevice, c->stream);
    if (cc) cudaMemcpyAsync(k->C, pkey + pPC + (size_t)k->lo[SEC_C] * 64, cc * 64, cudaMemcpyHostToDevice, c->stream);
    if (hc) cudaMemcpyAsync(k->H, pkey + pPH + (size_t)k->lo[SEC_H] * 64, hc * 64, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(k->fixed, pkey + 40, 448, cudaMemcpyHostToDevice, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
        pkey_release(k);
        return c->fail_cuda("proving key upload");
    }
    lap("points h2d");
    if (c->key_tables) {
        build_key_tables(c, k);
        lap("window tables");
        if (cudaMalloc(&k->fixed_tables, sizeof(FixedBaseTables)) == cudaSuccess) {
            k_build_fixed_tables<<<(D1_WINDOWS + D2_WINDOWS + 63) / 64, 64, 0, c->stream>>>(k->fixed, k->fixed_tables);
            if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
                cudaGetLastError();
                cudaFree(k->fixed_tables);
                k->fixed_tables = nullptr;
            }
        } else {
            cudaGetLastError();
            k->fixed_tables = nullptr;
        }
        lap("fixed tables");
    }
    *out = k;
    return WS_OK;
}

// ============================================================================ H pipeline + prover (device)
static bool ensure_twiddles(ws_ctx* c, int logM, cudaStream_t st) {
    TwiddleTable T;
    T.tw = (Fr*)c->tw_table;
    T.logM = c->tw_logM;
    bool ok = ntt_ensure_table(T, logM, st);
    c->tw_table = T.tw;
    c->tw_logM = T.logM;
    return ok;
}
static TwiddleTable twiddles(ws_ctx* c) {
    TwiddleTable T;
    T.tw = (Fr*)c->tw_table;
    T.logM = c->tw_logM;
    return T;
}

// h = coefficients n..2n-1 of A*B, out of Montgomery form; everything on `st`.
// work: (n_signals + 8n) field elements of scratch.
static int calc_h_dev(ws_ctx* c, const DevCsr& pa, const DevCsr& pb, const Fr* d_signals, uint32_t n_signals, int logn, Fr* d_h, Fr* work, cudaStream_t st) {
    const uint32_t n = 1u << logn;
    if (!ensure_twiddles(c, logn + 1, st)) return c->fail_cuda("twiddle table");
    TwiddleTable T = twiddles(c);
    Fr* sM = work;
    Fr* A = sM + n_signals;
    Fr* B = A + n;
    Fr* Ao = B + n;
    Fr* Bo = Ao + n;
    Fr* P = Bo + n;            // 2n
    Fr* scratch = P + 2 * (size_t)n;   // 2n
    c->timer.mark("h:begin", st);
    k_fr_to_mont<<<(n_signals + 127) / 128, 128, 0, st>>>(d_signals, sM, n_signals);
    k_csr_lc<<<(n + 127) / 128, 128, 0, st>>>(pa.row_start, pa.sig, pa.coef, sM, A, n);
    k_csr_lc<<<(n + 127) / 128, 128, 0, st>>>(pb.row_start, pb.sig, pb.coef, sM, B, n);
    c->timer.mark("h:lc", st);
    const int nv = c->ntt_variant;
    // A, B keep the evaluations on the even domain; Ao, Bo receive the coefficients (out-of-place inverse transform:
    // no copy), then become the evaluations on the odd coset in place (P is free until the product and serves as the
    // second scratch buffer an odd pass count needs)
    ntt_launch<NttBn128>(A, Ao, scratch, nullptr, logn, 0, 1, T, st, &c->launches, nv);
    ntt_launch<NttBn128>(B, Bo, scratch, nullptr, logn, 0, 1, T, st, &c->launches, nv);
    ntt_launch<NttBn128>(Ao, Ao, scratch, P, logn, 1, 0, T, st, &c->launches, nv);
    ntt_launch<NttBn128>(Bo, Bo, scratch, P, logn, 1, 0, T, st, &c->launches, nv);
    k_interleave_mul<<<(n + 127) / 128, 128, 0, st>>>(A, B, Ao, Bo, P, n);
    ntt_launch<NttBn128>(P, P, scratch, A, logn + 1, 0, 1, T, st, &c->launches, nv);     // A|B (2n, contiguous) are dead now
    c->launches += 5;      // to_mont, 2 x csr_lc, interleave_mul, from_mont
    k_fr_from_mont<<<(n + 127) / 128, 128, 0, st>>>(P + n, d_h, n);
    c->timer.mark("h:ntt", st);
    return cudaGetLastError() == cudaSuccess ? WS_OK : c->fail_cuda("calc_h");
}

// The five partial sums of this key's shard, as XYZZ records: A | B1 | C | H (128 bytes each) | B2 (256).
constexpr size_t PARTIALS_BYTES = 4 * sizeof(G1X) + sizeof(G2X);   // 768

struct PartialSlots {
    G1X *rA, *rB1, *rC, *rH;
    G2X* rB2;
    explicit PartialSlots(uint8_t* d_partials) {
        rA = (G1X*)d_partials;
        rB1 = rA + 1;
        rC = rB1 + 1;
        rH = rC + 1;
        rB2 = (G2X*)(rH + 1);
    }
};

// The four witness MSMs of this key's shard (A, B1, C over G1; B2 over G2) on side streams forked from `st` at this
// point, each with its own workspace -- the reference's promise graph (src/bn128.js:607-622), with streams instead of
// workers.  They do not depend on H.  prove_h_msm_dev (or prove_join_dev) joins them back into `st`.
static int prove_witness_msms_dev(ws_ctx* c, const ws_pkey* k, const uint8_t* d_witness, uint8_t* d_partials, cudaStream_t st) {
    if (!c->ensure_lanes()) return c->fail_cuda("prover streams");
    PartialSlots p(d_partials);
    cudaEventRecord(c->fork_ev, st);
    int rc = 0;
    for (int lane = 0; lane < 4 && !rc; lane++) {
        cudaStream_t ls = c->lanes[lane];
        cudaStreamWaitEvent(ls, c->fork_ev, 0);
        const int sec = lane == 0 ? SEC_B2 : lane == 1 ? SEC_A : lane == 2 ? SEC_B1 : SEC_C;      // the G2 MSM (the longest) first
        const uint8_t* sc = d_witness + 32 * ((size_t)k->lo[sec] + (sec == SEC_C ? (size_t)k->n_public + 1 : 0));
        void* out = sec == SEC_B2 ? (void*)p.rB2 : sec == SEC_A ? (void*)p.rA : sec == SEC_B1 ? (void*)p.rB1 : (void*)p.rC;
        if (sec == SEC_B2) rc = msm_g2_run(c, sc, k->msm_points(sec), k->cnt[sec], nullptr, out, ls, 8 + lane, k->tc[sec]);
        else rc = msm_g1_run(c, sc, k->msm_points(sec), k->cnt[sec], nullptr, out, ls, 8 + lane, k->tc[sec]);
        cudaEventRecord(c->lane_done[lane], ls);
    }
    if (rc) return rc;
    return cudaGetLastError() == cudaSuccess ? WS_OK : c->fail_cuda("prove: witness MSMs");
}

// h = coefficients n..2n-1 of A*B for the WHOLE domain (the reference's CALC_H task), from the whole witness, on `st`.
static int prove_h_dev(ws_ctx* c, const ws_pkey* k, const uint8_t* d_witness, Fr* d_h, cudaStream_t st) {
    const uint32_t nv = k->n_vars, n = k->domain;
    const size_t work_elems = (size_t)nv + 8 * (size_t)n + 64;
    Fr* work = (Fr*)c->arena(2, work_elems * 32);
    if (!work) return c->fail_cuda("prover workspace");
    return calc_h_dev(c, k->polA, k->polB, (const Fr*)d_witness, nv, k->logn, d_h, work, st);
}

// The H MSM of this key's shard: d_h_shard = h[h_lo .. h_lo + h_cnt), on `st`; then the side streams of the witness
// MSMs are joined into `st`, after which the 768-byte record is complete in stream order.
static int prove_h_msm_dev(ws_ctx* c, const ws_pkey* k, const Fr* d_h_shard, uint8_t* d_partials, cudaStream_t st) {
    PartialSlots p(d_partials);
    int rc = msm_g1_run(c, d_h_shard, k->msm_points(SEC_H), k->cnt[SEC_H], nullptr, p.rH, st, 1, k->tc[SEC_H]);
    if (rc) return rc;
    c->timer.mark("msm:H", st);
    for (int lane = 0; lane < 4; lane++) cudaStreamWaitEvent(st, c->lane_done[lane], 0);
    c->timer.mark("msm:A,B1,B2,C (side streams)", st);
    return cudaGetLastError() == cudaSuccess ? WS_OK : c->fail_cuda("prove: H MSM");
}

static int prove_partials_dev(ws_ctx* c, const ws_pkey* k, const uint8_t* d_witness, uint8_t* d_partials, cudaStream_t st) {
    Fr* h = (Fr*)c->arena(7, ((size_t)k->domain + 8) * 32);
    if (!h || !c->ensure_lanes()) return c->fail_cuda("prover workspace");
    c->timer.begin(st);
    int rc = prove_witness_msms_dev(c, k, d_witness, d_partials, st);
    if (rc) return rc;
    // the dependent chain (LC -> NTTs -> H MSM) optionally runs on the high-priority lane, forked from / joined to `st`
    cudaStream_t hs = c->h_priority ? c->lanes[5] : st;
    if (hs != st) cudaStreamWaitEvent(hs, c->fork_ev, 0);
    if ((rc = prove_h_dev(c, k, d_witness, h, hs))) return rc;
    if (hs != st) {
        cudaEventRecord(c->h_done, hs);
        cudaStreamWaitEvent(st, c->h_done, 0);
    }
    return prove_h_msm_dev(c, k, h + k->lo[SEC_H], d_partials, st);
}

// Starts the (r, s)-only products on a side stream; `st` must have d_rs ready at this point.
static int prove_pre_dev(ws_ctx* c, const ws_pkey* k, const uint8_t* d_rs, cudaStream_t st) {
    PreProducts* pre = (PreProducts*)c->arena(6, sizeof(PreProducts));
    if (!pre || !c->ensure_lanes()) return c->fail_cuda("prover workspace");
    cudaStream_t ls = c->lanes[4];
    cudaEventRecord(c->lane_done[5], st);
    cudaStreamWaitEvent(ls, c->lane_done[5], 0);
    c->launches += 1;
    k_assemble_pre<<<1, 128, 0, ls>>>(k->fixed, k->fixed_tables, d_rs, pre);
    cudaEventRecord(c->lane_done[4], ls);
    c->pre_pending = true;
    return cudaGetLastError() == cudaSuccess ? WS_OK : c->fail_cuda("assemble_pre");
}

// Adds up the `world` partial records (the exchange step of the sharded prover: the reference's serial
// g1m_add over worker results, src/bn128.js:376-380) -- one thread per MSM.
static __global__ void __launch_bounds__(32) k_sum_partials(const uint8_t* __restrict__ all, int world, uint8_t* __restrict__ out) {
    int i = threadIdx.x;
    if (i < 4) {
        G1X acc = G1X::inf();
        for (int r = 0; r < world; r++) acc = padd_call<G1Curve>(acc, *((const G1X*)(all + (size_t)r * PARTIALS_BYTES) + i));
        ((G1X*)out)[i] = acc;
    } else if (i == 4) {
        G2X acc = G2X::inf();
        for (int r = 0; r < world; r++) acc = padd_call<G2Curve>(acc, *(const G2X*)(all + (size_t)r * PARTIALS_BYTES + 4 * sizeof(G1X)));
        *(G2X*)(out + 4 * sizeof(G1X)) = acc;
    }
}

// d_all: world x 768 bytes of partial records; d_out: 96 (pi_a) + 192 (pi_b) + 96 (pi_c) bytes, affine, out of Montgomery form.
static int prove_finish_dev(ws_ctx* c, const ws_pkey* k, const uint8_t* d_all, int world, const uint8_t* d_rs, uint8_t* d_out, cudaStream_t st) {
    uint8_t* sum = (uint8_t*)c->arena(4, 1024);
    if (!sum) return c->fail_cuda("prover workspace");
    const uint8_t* src = d_all;
    if (world > 1) {
        k_sum_partials<<<1, 32, 0, st>>>(d_all, world, sum);
        src = sum;
    }
    if (!c->pre_pending) {
        int rc = prove_pre_dev(c, k, d_rs, st);
        if (rc) return rc;
    }
    cudaStreamWaitEvent(st, c->lane_done[4], 0);
    c->pre_pending = false;
    AssembleArgs g;
    g.A = (const G1X*)src;
    g.B1 = g.A + 1;
    g.C = g.A + 2;
    g.H = g.A + 3;
    g.B2 = (const G2X*)(g.A + 4);
    g.fx = k->fixed;
    g.rs = d_rs;
    g.pi_a = (Jac<G1Curve>*)d_out;
    g.pi_b = (Jac<G2Curve>*)(d_out + 96);
    g.pi_c = (Jac<G1Curve>*)(d_out + 288);
    c->launches += 1 + (world > 1 ? 1 : 0);
    k_assemble_post<<<1, 128, 0, st>>>(g, (const PreProducts*)c->arena(6, sizeof(PreProducts)));
    c->timer.mark("assemble", st);
    return cudaGetLastError() == cudaSuccess ? WS_OK : c->fail_cuda("prove");
}

static int prove_dev(ws_ctx* c, const ws_pkey* k, const uint8_t* d_witness, const uint8_t* d_rs, uint8_t* d_out, cudaStream_t st) {
    if (k->world != 1) return c->fail(WS_E_ARG, "sharded key: use the partials / finish entry points");
    uint8_t* part = (uint8_t*)c->arena(5, 1024);
    if (!part) return c->fail_cuda("prover workspace");
    int rc = prove_pre_dev(c, k, d_rs, st);       // overlaps everything below
    if (rc) return rc;
    if ((rc = prove_partials_dev(c, k, d_witness, part, st))) return rc;
    return prove_finish_dev(c, k, part, 1, d_rs, d_out, st);
}

// The blocking host-buffer call in two halves, so that the stateless entry point can fingerprint the caller's key
// buffer on the host while the GPU already works.
//   prove_host_enqueue  stages the witness to the device (pinned ring for pageable memory) and enqueues the proof
//   prove_host_collect  waits and copies the 384 result bytes out
static int prove_host_enqueue(ws_ctx* c, const ws_pkey* k, const uint8_t* witness, size_t wlen, const uint8_t* r, const uint8_t* s) {
    if (wlen < (size_t)k->n_vars * 32) return c->fail(WS_E_ARG, "witness shorter than nVars * 32 bytes");
    cudaSetDevice(c->device);
    uint8_t rs[64];
    // r, s: 32 random bytes each (src/bn128.js:642-664) unless injected
    if (r) memcpy(rs, r, 32);
    else if (getrandom(rs, 32, 0) != 32) return c->fail(WS_E_ARG, "getrandom failed");
    if (s) memcpy(rs + 32, s, 32);
    else if (getrandom(rs + 32, 32, 0) != 32) return c->fail(WS_E_ARG, "getrandom failed");
    size_t wbytes = (size_t)k->n_vars * 32;
    uint8_t* d = (uint8_t*)c->arena(3, align256(wbytes) + 1024);
    if (!d) return c->fail_cuda("witness buffer");
    uint8_t* d_rs = d + align256(wbytes);
    uint8_t* d_out = d_rs + 256;
    if (c->stager.upload(d, witness, wbytes, c->stream, c->device) != cudaSuccess) return c->fail_cuda("witness upload");
    if (cudaMemcpyAsync(d_rs, rs, 64, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return c->fail_cuda("r, s upload");   // pageable: staged before return
    return prove_dev(c, k, d, d_rs, d_out, c->stream);
}

static int prove_host_collect(ws_ctx* c, const ws_pkey* k, uint8_t* pi_a, uint8_t* pi_b, uint8_t* pi_c) {
    size_t wbytes = (size_t)k->n_vars * 32;
    uint8_t* d = (uint8_t*)c->arena(3, align256(wbytes) + 1024);
    uint8_t* d_out = d + align256(wbytes) + 256;
    uint8_t outb[384];
    if (cudaMemcpyAsync(outb, d_out, 384, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return c->fail_cuda("proof download");
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return c->fail_cuda("prove");
    memcpy(pi_a, outb, 96);
    memcpy(pi_b, outb + 96, 192);
    memcpy(pi_c, outb + 288, 96);
    return WS_OK;
}

static int prove_host(ws_ctx* c, const ws_pkey* k, const uint8_t* witness, size_t wlen, const uint8_t* r, const uint8_t* s, uint8_t* pi_a, uint8_t* pi_b,
                      uint8_t* pi_c) {
    int rc = prove_host_enqueue(c, k, witness, wlen, r, s);
    if (rc) return rc;
    return prove_host_collect(c, k, pi_a, pi_b, pi_c);
}

// ============================================================================ C ABI
extern "C" {

WS_API int ws_bn128_fr_fft_dev(ws_ctx* c, void* d_px, size_t n, int odd, int inverse, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !d_px) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    int logn = n <= 0xFFFFFFFFull ? log2_exact((uint32_t)n) : -1;
    if (logn < 0 || logn > 28) return c->fail(WS_E_SIZE, "fft size must be a power of two <= 2^28");
    cudaSetDevice(c->device);
    // BN128 Fr has 2-adicity 28: the odd (coset) transform needs w_{2n}, which does not exist for n = 2^28
    // (the reference's table ends at w[28] as well, src/build_fft.js:29-57)
    if (odd && logn == 28) return c->fail(WS_E_SIZE, "odd fft of size 2^28 needs a 2^29-th root of unity");
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    if (!ensure_twiddles(c, logn + (odd ? 1 : 0), st)) return c->fail_cuda("twiddle table");
    Fr* scratch = (Fr*)c->arena(2, 2 * (sizeof(Fr) << logn));       // two buffers: an odd pass count needs both
    if (!scratch) return c->fail_cuda("fft scratch");
    WsStreamScope chain(c, st);
    c->timer.begin(st);
    cudaError_t e = ntt_launch<NttBn128>((const Fr*)d_px, (Fr*)d_px, scratch, scratch + ((size_t)1 << logn), logn, odd != 0, inverse != 0, twiddles(c), st,
                                         &c->launches, c->ntt_variant);
    c->timer.mark("ntt", st);
    return e == cudaSuccess ? WS_OK : c->fail_cuda("ntt");
}

static int fft_host(ws_ctx* c, uint8_t* px, size_t n, int odd, int inverse) {
    WsLock ws_lock_(c);
    if (!c || !px) return c ? c->fail(WS_E_ARG, "null buffer") : WS_E_ARG;
    int logn = n <= 0xFFFFFFFFull ? log2_exact((uint32_t)n) : -1;
    if (logn < 0 || logn > 28) return c->fail(WS_E_SIZE, "fft size must be a power of two <= 2^28");
    cudaSetDevice(c->device);
    uint8_t* d = (uint8_t*)c->arena(3, n * 32);
    if (!d) return c->fail_cuda("fft buffer");
    cudaMemcpyAsync(d, px, n * 32, cudaMemcpyHostToDevice, c->stream);
    int rc = ws_bn128_fr_fft_dev(c, d, n, odd, inverse, c->stream);
    if (rc) return rc;
    cudaMemcpyAsync(px, d, n * 32, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return c->fail_cuda("fft");
    return WS_OK;
}
WS_API int ws_bn128_fr_fft(ws_ctx* c, uint8_t* px, size_t n, int odd) { return fft_host(c, px, n, odd, 0); }
WS_API int ws_bn128_fr_ifft(ws_ctx* c, uint8_t* px, size_t n, int odd) { return fft_host(c, px, n, odd, 1); }

WS_API int ws_set_h_priority(ws_ctx* c, int on) {
    WsLock ws_lock_(c);
    if (!c) return WS_E_ARG;
    c->h_priority = on != 0;
    return WS_OK;
}
WS_API int ws_set_key_tables(ws_ctx* c, int on) {
    WsLock ws_lock_(c);
    if (!c) return WS_E_ARG;
    c->key_tables = on != 0;
    return WS_OK;
}
WS_API int ws_bn128_pkey_load(ws_ctx* c, const uint8_t* pkey, size_t len, ws_pkey** out) {
    WsLock ws_lock_(c);
    if (!c || !pkey || !out) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    *out = nullptr;
    return pkey_load(c, pkey, len, out);
}
WS_API int ws_bn128_pkey_load_shard(ws_ctx* c, const uint8_t* pkey, size_t len, int rank, int world, ws_pkey** out) {
    WsLock ws_lock_(c);
    if (!c || !pkey || !out) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    *out = nullptr;
    return pkey_load(c, pkey, len, out, rank, world);
}
WS_API int ws_bn128_groth16_partials_dev(ws_ctx* c, const ws_pkey* k, const void* d_witness, void* d_partials, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !k || !d_witness || !d_partials) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    if (k->ctx != c) return c->fail(WS_E_ARG, "proving key belongs to another context");
    if (((uintptr_t)d_witness | (uintptr_t)d_partials) & 15) return c->fail(WS_E_ARG, "device pointers must be 16-byte aligned");
    cudaSetDevice(c->device);
    WsStreamScope chain(c, stream ? (cudaStream_t)stream : c->stream);
    return prove_partials_dev(c, k, (const uint8_t*)d_witness, (uint8_t*)d_partials, stream ? (cudaStream_t)stream : c->stream);
}
WS_API int ws_bn128_groth16_witness_msms_dev(ws_ctx* c, const ws_pkey* k, const void* d_witness, void* d_partials, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !k || !d_witness || !d_partials) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    if (k->ctx != c) return c->fail(WS_E_ARG, "proving key belongs to another context");
    if (((uintptr_t)d_witness | (uintptr_t)d_partials) & 15) return c->fail(WS_E_ARG, "device pointers must be 16-byte aligned");
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    c->enter_stream(st);            // the matching leave_stream is in ws_bn128_groth16_h_msm_dev, which joins the side streams
    c->timer.begin(st);
    return prove_witness_msms_dev(c, k, (const uint8_t*)d_witness, (uint8_t*)d_partials, st);
}
WS_API int ws_bn128_groth16_h_dev(ws_ctx* c, const ws_pkey* k, const void* d_witness, void* d_h, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !k || !d_witness || !d_h) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    if (k->ctx != c) return c->fail(WS_E_ARG, "proving key belongs to another context");
    if (((uintptr_t)d_witness | (uintptr_t)d_h) & 15) return c->fail(WS_E_ARG, "device pointers must be 16-byte aligned");
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    return prove_h_dev(c, k, (const uint8_t*)d_witness, (Fr*)d_h, st);
}
WS_API int ws_bn128_groth16_h_msm_dev(ws_ctx* c, const ws_pkey* k, const void* d_h_shard, void* d_partials, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !k || !d_partials || (k->cnt[SEC_H] && !d_h_shard)) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    if (k->ctx != c) return c->fail(WS_E_ARG, "proving key belongs to another context");
    if (((uintptr_t)d_h_shard | (uintptr_t)d_partials) & 15) return c->fail(WS_E_ARG, "device pointers must be 16-byte aligned");
    if (!c->lanes_ready) return c->fail(WS_E_ARG, "ws_bn128_groth16_witness_msms_dev must be called first");
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc = prove_h_msm_dev(c, k, (const Fr*)d_h_shard, (uint8_t*)d_partials, st);
    c->leave_stream(st);
    return rc;
}
WS_API int ws_bn128_pkey_shard_info(const ws_pkey* k, uint32_t range[6]) {
    if (!k || !range) return WS_E_ARG;
    // the hull of the A / B1 / B2 ranges (identical for the uniform shard), the C range, the h range
    uint32_t vlo = 0xFFFFFFFFu, vhi = 0;
    for (int sec = SEC_A; sec <= SEC_B2; sec++)
        if (k->cnt[sec]) {
            vlo = k->lo[sec] < vlo ? k->lo[sec] : vlo;
            vhi = k->lo[sec] + k->cnt[sec] > vhi ? k->lo[sec] + k->cnt[sec] : vhi;
        }
    range[0] = vhi ? vlo : 0;
    range[1] = vhi ? vhi - vlo : 0;
    range[2] = k->n_public + 1 + k->lo[SEC_C];
    range[3] = k->cnt[SEC_C];
    range[4] = k->lo[SEC_H];
    range[5] = k->cnt[SEC_H];
    return WS_OK;
}
WS_API int ws_bn128_pkey_load_ranges(ws_ctx* c, const uint8_t* pkey, size_t len, const uint32_t ranges[10], ws_pkey** out) {
    WsLock ws_lock_(c);
    if (!c || !pkey || !out || !ranges) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    *out = nullptr;
    int rc = pkey_load(c, pkey, len, out, 0, 1, ranges);
    if (!rc) (*out)->world = 0;          // a partial key: only the partials / finish entry points accept it
    return rc;
}
WS_API int ws_bn128_groth16_prepare_dev(ws_ctx* c, const ws_pkey* k, const void* d_rs, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !k || !d_rs) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    if (k->ctx != c) return c->fail(WS_E_ARG, "proving key belongs to another context");
    cudaSetDevice(c->device);
    WsStreamScope chain(c, stream ? (cudaStream_t)stream : c->stream);
    return prove_pre_dev(c, k, (const uint8_t*)d_rs, stream ? (cudaStream_t)stream : c->stream);
}
WS_API int ws_bn128_groth16_finish_dev(ws_ctx* c, const ws_pkey* k, const void* d_all_partials, int world, const void* d_rs, void* d_out, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !k || !d_all_partials || !d_rs || !d_out || world < 1) return c ? c->fail(WS_E_ARG, "bad argument") : WS_E_ARG;
    if (k->ctx != c) return c->fail(WS_E_ARG, "proving key belongs to another context");
    cudaSetDevice(c->device);
    WsStreamScope chain(c, stream ? (cudaStream_t)stream : c->stream);
    return prove_finish_dev(c, k, (const uint8_t*)d_all_partials, world, (const uint8_t*)d_rs, (uint8_t*)d_out, stream ? (cudaStream_t)stream : c->stream);
}
WS_API void ws_bn128_pkey_free(ws_pkey* k) {
    if (!k) return;
    cudaSetDevice(k->ctx->device);
    cudaStreamSynchronize(k->ctx->stream);
    pkey_release(k);
}
WS_API int ws_bn128_pkey_info(const ws_pkey* k, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_size) {
    if (!k) return WS_E_ARG;
    if (n_vars) *n_vars = k->n_vars;
    if (n_public) *n_public = k->n_public;
    if (domain_size) *domain_size = k->domain;
    return WS_OK;
}

WS_API int ws_bn128_calc_h(ws_ctx* c, const uint8_t* signals, const uint8_t* polsA, size_t lenA, const uint8_t* polsB, size_t lenB, uint32_t n_signals,
                           uint32_t domain_size, uint8_t* out) {
    WsLock ws_lock_(c);
    if (!c || !signals || !polsA || !polsB || !out) return c ? c->fail(WS_E_ARG, "null buffer") : WS_E_ARG;
    int logn = log2_exact(domain_size);
    if (logn < 0 || logn > 27) return c->fail(WS_E_SIZE, "domainSize must be a power of two <= 2^27");
    cudaSetDevice(c->device);
    std::vector<uint32_t> rs, sg;
    std::vector<uint8_t> cf;
    DevCsr da, db;
    int rc = WS_OK;
    if (!build_csr(polsA, lenA, n_signals, domain_size, rs, sg, cf) || !upload_csr(rs, sg, cf, da, c->stream)) rc = c->fail(WS_E_ARG, "polsA malformed");
    if (!rc && (!build_csr(polsB, lenB, n_signals, domain_size, rs, sg, cf) || !upload_csr(rs, sg, cf, db, c->stream))) rc = c->fail(WS_E_ARG, "polsB malformed");
    if (!rc) {
        size_t n = domain_size;
        uint8_t* base = (uint8_t*)c->arena(2, ((size_t)2 * n_signals + 9 * n + 64) * 32);
        if (!base) rc = c->fail_cuda("calc_h workspace");
        else {
            Fr* d_sig = (Fr*)base;
            Fr* h = d_sig + n_signals;
            Fr* work = h + n;
            cudaMemcpyAsync(d_sig, signals, (size_t)n_signals * 32, cudaMemcpyHostToDevice, c->stream);
            c->timer.begin(c->stream);
            rc = calc_h_dev(c, da, db, d_sig, n_signals, logn, h, work, c->stream);
            if (!rc) {
                cudaMemcpyAsync(out, h, n * 32, cudaMemcpyDeviceToHost, c->stream);
                if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = c->fail_cuda("calc_h");
            }
        }
    }
    cudaStreamSynchronize(c->stream);
    da.free_all();
    db.free_all();
    return rc;
}

WS_API int ws_bn128_groth16_prove_key(ws_ctx* c, const ws_pkey* k, const uint8_t* witness, size_t wlen, const uint8_t* r, const uint8_t* s, uint8_t pi_a[96],
                                      uint8_t pi_b[192], uint8_t pi_c[96]) {
    WsLock ws_lock_(c);
    if (!c || !k || !witness || !pi_a || !pi_b || !pi_c) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    if (k->ctx != c) return c->fail(WS_E_ARG, "proving key belongs to another context");
    return prove_host(c, k, witness, wlen, r, s, pi_a, pi_b, pi_c);
}

// 64-bit content fingerprint of a key buffer (4 independent multiply-rotate lanes over 8-byte words, ~10 GB/s):
// guards the one-entry key cache of the stateless call against a different key at the same address / length.
static unsigned long long key_fingerprint_1(const uint8_t* p, size_t n) {
    unsigned long long h[4] = {0x243F6A8885A308D3ull ^ n, 0x13198A2E03707344ull, 0xA4093822299F31D0ull, 0x082EFA98EC4E6C89ull};
    const unsigned long long M = 0x9E3779B97F4A7C15ull;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        for (int k = 0; k < 4; k++) {
            unsigned long long w;
            memcpy(&w, p + i + 8 * k, 8);
            h[k] = (h[k] ^ w) * M;
            h[k] = (h[k] << 29) | (h[k] >> 35);
        }
    }
    for (; i < n; i++) h[i & 3] = (h[i & 3] ^ p[i]) * M;
    unsigned long long r = h[0];
    for (int k = 1; k < 4; k++) r = (r ^ (h[k] + (r << 6) + (r >> 2))) * M;
    return r ^ (r >> 31);
}

// large keys (hundreds of MB at 2^20 constraints) are hashed in 4 MiB-aligned slices by a few host threads
static unsigned long long key_fingerprint(const uint8_t* p, size_t n) {
    const size_t kMin = (size_t)8 << 20;
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::min<size_t>(std::min<unsigned>(hw ? hw : 1, 8), n / kMin);
    if (nt <= 1) return key_fingerprint_1(p, n);
    std::vector<unsigned long long> part(nt);
    std::vector<std::thread> th;
    size_t per = ((n + nt - 1) / nt + 31) & ~(size_t)31;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t] {
            size_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
            part[t] = key_fingerprint_1(p + lo, hi - lo);
        });
    for (auto& x : th) x.join();
    return key_fingerprint_1((const uint8_t*)part.data(), part.size() * sizeof(part[0])) ^ n;
}

static void cached_key_free(void* k) { pkey_release((ws_pkey*)k); }

WS_API int ws_set_key_cache(ws_ctx* c, int on) {
    WsLock ws_lock_(c);
    if (!c) return WS_E_ARG;
    c->key_cache = on != 0;
    if (!on && c->cached_key) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        pkey_release((ws_pkey*)c->cached_key);
        c->cached_key = nullptr;
    }
    return WS_OK;
}

WS_API int ws_bn128_groth16_prove(ws_ctx* c, const uint8_t* witness, size_t wlen, const uint8_t* pkey, size_t plen, const uint8_t* r, const uint8_t* s,
                                  uint8_t pi_a[96], uint8_t pi_b[192], uint8_t pi_c[96]) {
    WsLock ws_lock_(c);
    if (!c || !witness || !pkey || !pi_a || !pi_b || !pi_c) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    if (!c->key_cache) {
        ws_pkey* k = nullptr;
        int rc = pkey_load(c, pkey, plen, &k);
        if (rc) return rc;
        rc = prove_host(c, k, witness, wlen, r, s, pi_a, pi_b, pi_c);
        cudaStreamSynchronize(c->stream);
        pkey_release(k);
        return rc;
    }
    // The reference is handed the same proving key on every call and re-slices / re-posts it each time
    // (src/bn128.js:581-604).  Here the parsed, uploaded key (with its window tables) of the previous call is
    // kept and reused when the caller's buffer has the same length and content fingerprint.
    // The whole buffer is fingerprinted on every call (a sampled hash could miss an in-place edit), but not on the
    // critical path: when an entry of the same length exists, the proof is enqueued with it FIRST and the host threads
    // hash the caller's buffer while the GPU works; only a mismatch (a different key) discards that proof, reloads
    // and proves again.  r and s are drawn once so that both attempts would use the same values.
    uint8_t rs[64];
    if (r) memcpy(rs, r, 32);
    else if (getrandom(rs, 32, 0) != 32) return c->fail(WS_E_ARG, "getrandom failed");
    if (s) memcpy(rs + 32, s, 32);
    else if (getrandom(rs + 32, 32, 0) != 32) return c->fail(WS_E_ARG, "getrandom failed");
    bool optimistic = c->cached_key && c->cached_len == plen;
    if (optimistic && prove_host_enqueue(c, (const ws_pkey*)c->cached_key, witness, wlen, rs, rs + 32) != WS_OK) {
        cudaStreamSynchronize(c->stream);       // e.g. a same-length key with another nVars: decide after the fingerprint
        cudaGetLastError();
        optimistic = false;
    }
    unsigned long long fp = key_fingerprint(pkey, plen);
    if (optimistic && c->cached_fp == fp) return prove_host_collect(c, (const ws_pkey*)c->cached_key, pi_a, pi_b, pi_c);
    cudaSetDevice(c->device);
    if (c->cached_key) {
        cudaStreamSynchronize(c->stream);
        pkey_release((ws_pkey*)c->cached_key);
        c->cached_key = nullptr;
    }
    ws_pkey* k = nullptr;
    int rc = pkey_load(c, pkey, plen, &k);
    if (rc) return rc;
    c->cached_key = k;
    c->cached_fp = fp;
    c->cached_len = plen;
    c->cached_key_free = cached_key_free;
    return prove_host(c, k, witness, wlen, rs, rs + 32, pi_a, pi_b, pi_c);
}

WS_API int ws_bn128_groth16_prove_dev(ws_ctx* c, const ws_pkey* k, const void* d_witness, const void* d_rs, void* d_out, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !k || !d_witness || !d_rs || !d_out) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    if (k->ctx != c) return c->fail(WS_E_ARG, "proving key belongs to another context");
    if (((uintptr_t)d_witness | (uintptr_t)d_out) & 15) return c->fail(WS_E_ARG, "device pointers must be 16-byte aligned");
    cudaSetDevice(c->device);
    WsStreamScope chain(c, stream ? (cudaStream_t)stream : c->stream);
    return prove_dev(c, k, (const uint8_t*)d_witness, (const uint8_t*)d_rs, (uint8_t*)d_out, stream ? (cudaStream_t)stream : c->stream);
}

}  // extern "C"
