// msm_core.cuh -- per-thread bodies of the multi-scalar-multiplication pipeline.
//
// Replaces g1m_multiexp2 / g2m_multiexp of the reference
//   /root/reference/src/build_multiexp.js:498-580 (multiexp), :651-744 (multiexp2),
//   called with w = 7 from /root/reference/src/bn128.js:109,121
// with a bucket (Pippenger) method.  What is kept from the reference, because it decides the
// result: scalars are RAW 256-bit little-endian integers, never reduced mod r
// (build_multiexp.js:30-155 slices all 256 bits); an affine key point whose x coordinate is
// zero is the point at infinity (:335-349); the sum is a group element, compared after
// g?m_affine.  What is free and changed: window size, bucketing, summation order.
//
// Pipeline (one launch each, all on one stream, no host round trip):
//   digits     signed c-bit digits of every scalar (carry into a possible extra window),
//              bucket histogram with global atomics
//   scan       exclusive scan of the histogram -> bucket offsets; second scan of
//              ceil(len/T) -> task offsets
//   scatter    counting sort of (point index, sign) into bucket order
//   accumulate one thread per task (<= T entries of one bucket): gather affine points,
//              8M+2S mixed additions into an XYZZ register accumulator
//   combine    buckets split over several tasks: one warp per bucket, shuffle tree
//   segreduce  per segment of L consecutive buckets: running-sum pair (A, T)
//   bitsum     per window: sum of A's and the log2(S) bit-plane sums of the T's
//   final      per window Horner over bit planes, then Horner over windows, affine
//
// The bodies below are HD so that tests/host/host_msm.cpp can run the identical index
// logic sequentially on a CPU (no GPU in the build container); the cooperative steps
// (scan, warp combine, block bit-sum) have their parallel versions in msm.cuh.
#pragma once
#include "ec.cuh"

namespace ws {

constexpr int MSM_MAX_LEVELS = 8;

struct MsmPlan {
    uint32_t n;        // number of (scalar, point) pairs
    int c;             // window bits, 2..22
    int sbits;         // scalar width in bits (256 for BN128, 768 for MNT6753)
    int ndig;          // digit windows incl. the carry window: ceil((sbits + 1) / c)
    int table;         // 1: the point array is a window table T[w * n + i] = 2^(c w) * P_i (resident keys), all
                       //    digit windows then share ONE bucket set and no doubling is ever needed
    int nwin;          // bucket windows: ndig, or 1 with a table
    uint32_t B;        // buckets per window = 2^(c-1); bucket b holds digit magnitude b+1
    uint32_t T;        // max entries per accumulate task
    uint32_t L;        // buckets per reduce segment (power of two, <= B)
    uint32_t S;        // segments per window = B / L
    int logS;
    // reduction tree above the segments: level 1 = the S segments; level l >= 2 groups G[l] items of level l-1
    // into Sl[l] groups per window, until at most 32 remain
    int nlev;                       // number of levels (>= 1)
    uint32_t G[MSM_MAX_LEVELS + 1]; // G[1] = L
    uint32_t Sl[MSM_MAX_LEVELS + 1];// Sl[1] = S
    uint32_t offT[MSM_MAX_LEVELS + 1];   // element offsets into the level pool: T array of level l
    uint32_t offC[MSM_MAX_LEVELS + 1];   // carried arrays C[0..l-1] of level l (contiguous, Sl[l] * nwin elements each)
    uint32_t pool_elems;
    int logSK;                      // log2 of the top level's group count
    uint32_t nbuckets; // nwin * B
    uint32_t max_tasks;
};

HD MsmPlan msm_make_plan(uint32_t n, int c, uint32_t T, uint32_t L = 0, int sbits = 256, int table = 0) {
    MsmPlan p;
    p.n = n;
    p.c = c;
    p.sbits = sbits;
    p.ndig = (sbits + 1 + c - 1) / c;
    p.table = table;
    p.nwin = table ? 1 : p.ndig;
    p.B = 1u << (c - 1);
    p.T = T;
    if (L == 0) {   // enough segments to occupy the part (~64K threads), at most 16 buckets each
        L = 16;
        while (L > 2 && (uint64_t)p.nwin * (p.B / L) < 65536) L >>= 1;
    }
    p.L = p.B < L ? p.B : L;
    p.S = p.B / p.L;
    p.logS = 0;
    while ((1u << p.logS) < p.S) p.logS++;
    p.nlev = 1;
    p.G[0] = p.Sl[0] = p.offT[0] = p.offC[0] = 0;
    p.G[1] = p.L;
    p.Sl[1] = p.S;
    while (p.Sl[p.nlev] > 32 && p.nlev < MSM_MAX_LEVELS) {
        uint32_t g = p.Sl[p.nlev] >= 8 ? 8 : p.Sl[p.nlev];
        p.nlev++;
        p.G[p.nlev] = g;
        p.Sl[p.nlev] = p.Sl[p.nlev - 1] / g;
    }
    for (int l = p.nlev + 1; l <= MSM_MAX_LEVELS; l++) p.G[l] = p.Sl[l] = p.offT[l] = p.offC[l] = 0;
    uint32_t off = 0;
    for (int l = 1; l <= p.nlev; l++) {
        uint32_t per = p.Sl[l] * (uint32_t)p.nwin;
        p.offT[l] = off;
        off += per;
        p.offC[l] = off;
        off += per * (uint32_t)l;
    }
    p.pool_elems = off;
    p.logSK = 0;
    while ((1u << p.logSK) < p.Sl[p.nlev]) p.logSK++;
    p.nbuckets = (uint32_t)p.nwin * p.B;
    uint64_t entries = (uint64_t)p.ndig * n;
    p.max_tasks = (uint32_t)(entries / T) + p.nbuckets + 1;
    return p;
}

HD uint32_t atomic_add_u32(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    uint32_t o = *p;
    *p = o + v;
    return o;
#endif
}

// ---- digits: thread i ------------------------------------------------------------------
// scalar_limbs: the SL u32 limbs (raw, LE) of scalar i.  is_inf: point i is infinity (x == 0).
// digits[w * n + i] = signed digit in [-2^(c-1), 2^(c-1)), counts[w' * B + |d| - 1]++  (w' = w, or 0 with a table).
template <int SL>
HD void msm_digits_body(const MsmPlan& p, uint32_t i, const uint32_t* scalar_limbs, bool is_inf,
                        int32_t* digits, uint32_t* counts) {
    uint32_t s[SL + 1];
#pragma unroll
    for (int k = 0; k < SL; k++) s[k] = is_inf ? 0u : scalar_limbs[k];
    s[SL] = 0;
    const uint32_t mask = (1u << p.c) - 1u;
    const uint32_t half = 1u << (p.c - 1);
    uint32_t carry = 0;
    for (int w = 0; w < p.ndig; w++) {
        int bit = w * p.c;
        uint32_t raw = 0;
        if (bit < 32 * SL) {
            int limb = bit >> 5, sh = bit & 31;
            uint64_t two = (uint64_t)s[limb] | ((uint64_t)s[limb + 1] << 32);
            raw = (uint32_t)(two >> sh) & mask;
        }
        uint32_t d = raw + carry;            // 0 .. 2^c
        int32_t sd;
        if (d >= half && w != p.ndig - 1) {   // the top window keeps its sign: nothing above it takes a carry
            sd = (int32_t)d - (int32_t)(mask + 1u);   // -2^(c-1) .. 0
            carry = 1;
        } else {
            sd = (int32_t)d;
            carry = 0;
        }
        digits[(size_t)w * p.n + i] = sd;
        if (sd != 0) {
            uint32_t mag = sd < 0 ? (uint32_t)(-sd) : (uint32_t)sd;
            atomic_add_u32(&counts[(size_t)(p.table ? 0 : w) * p.B + (mag - 1)], 1u);
        }
    }
}

// ---- scatter: thread i -----------------------------------------------------------------
// cursor[] starts as a copy of the histogram and is counted down; entries[] receives
// (point index | sign << 31) grouped by bucket (order inside a bucket is irrelevant to the sum).
HD void msm_scatter_body(const MsmPlan& p, uint32_t i, const int32_t* digits, const uint32_t* bstart,
                         uint32_t* cursor, uint32_t* entries) {
    for (int w = 0; w < p.ndig; w++) {
        int32_t sd = digits[(size_t)w * p.n + i];
        if (sd == 0) continue;
        uint32_t neg = sd < 0;
        uint32_t mag = neg ? (uint32_t)(-sd) : (uint32_t)sd;
        size_t b = (size_t)(p.table ? 0 : w) * p.B + (mag - 1);
        uint32_t k = atomic_add_u32(&cursor[b], 0xFFFFFFFFu) - 1u;   // counts down to 0
        entries[bstart[b] + k] = (p.table ? (uint32_t)w * p.n + i : i) | (neg << 31);   // index into the point / table array
    }
}

// ---- task table: thread b (one per bucket) ----------------------------------------------
HD uint32_t msm_ntasks(const MsmPlan& p, uint32_t len) { return (len + p.T - 1) / p.T; }

// Writes the bucket's task descriptors, counts task lengths (for the length sort that keeps the
// lanes of a warp in step) and files multi-task buckets for the combine step:
//   heavy[0 .. ) buckets with 2..32 tasks (one warp each), heavy[nbuckets .. ) buckets with more
//   (one block each); heavy_count[0], heavy_count[1] are the two list lengths.
HD void msm_filltasks_body(const MsmPlan& p, uint32_t b, const uint32_t* bstart, const uint32_t* tstart,
                           uint32_t* task_bucket, uint32_t* len_hist, uint32_t* heavy, uint32_t* heavy_count) {
    uint32_t t0 = tstart[b], t1 = tstart[b + 1];
    uint32_t blen = bstart[b + 1] - bstart[b];
    for (uint32_t t = t0; t < t1; t++) {
        task_bucket[t] = b;
        uint32_t rem = blen - (t - t0) * p.T;
        atomic_add_u32(&len_hist[rem < p.T ? rem : p.T], 1u);
    }
    uint32_t m = t1 - t0;
    if (m > 16) heavy[p.nbuckets + atomic_add_u32(heavy_count + 1, 1u)] = b;
    else if (m > 1) heavy[atomic_add_u32(heavy_count, 1u)] = b;
}

// Start offsets of each task length in the execution order: longest first.  One thread.
HD void msm_task_offsets_body(const MsmPlan& p, const uint32_t* len_hist, uint32_t* len_cursor) {
    uint32_t run = 0;
    for (uint32_t l = p.T; l >= 1; l--) {
        len_cursor[l] = run;
        run += len_hist[l];
    }
    len_cursor[0] = run;
}

// thread t: place task t in the execution order
HD void msm_task_order_body(const MsmPlan& p, uint32_t t, const uint32_t* bstart, const uint32_t* tstart,
                            const uint32_t* task_bucket, uint32_t* len_cursor, uint32_t* order) {
    uint32_t b = task_bucket[t];
    uint32_t rem = bstart[b + 1] - bstart[b] - (t - tstart[b]) * p.T;
    uint32_t len = rem < p.T ? rem : p.T;
    order[atomic_add_u32(&len_cursor[len], 1u)] = t;
}

// ---- accumulate: thread t (one per task) ------------------------------------------------
template <class C>
HD XYZZ<C> msm_accumulate_body(const MsmPlan& p, uint32_t t, const uint32_t* bstart, const uint32_t* tstart,
                               const uint32_t* task_bucket, const uint32_t* entries, const Affine<C>* points) {
    uint32_t b = task_bucket[t];
    uint32_t j = t - tstart[b];
    uint32_t lo = bstart[b] + j * p.T;
    uint32_t hi = bstart[b + 1];
    if (hi > lo + p.T) hi = lo + p.T;
    XYZZ<C> acc = XYZZ<C>::inf();
    for (uint32_t k = lo; k < hi; k++) {
        uint32_t e = entries[k];
        Affine<C> q = points[e & 0x7FFFFFFFu];
        if (e >> 31) q.y = neg(q.y);
        acc = pmadd<C>(acc, q);
    }
    return acc;
}

// ---- bucket value seen by the reduce phase ------------------------------------------------
template <class C>
HD XYZZ<C> msm_bucket_value(const uint32_t* tstart, const XYZZ<C>* partials, size_t b) {
    uint32_t t0 = tstart[b];
    if (tstart[b + 1] == t0) return XYZZ<C>::inf();
    return partials[t0];
}

// ---- segreduce: thread (w, s) ------------------------------------------------------------
// Over buckets b = s*L .. s*L+L-1 of window w (weights b+1):
//   T = sum S_b ;  A = sum (b - s*L + 1) * S_b      (running-sum trick, 2 adds per bucket)
template <class C>
HD void msm_segreduce_body(const MsmPlan& p, uint32_t w, uint32_t s, const uint32_t* tstart,
                           const XYZZ<C>* partials, XYZZ<C>* pool) {
    XYZZ<C>* segA = pool + p.offC[1];      // carried array 0 of level 1
    XYZZ<C>* segT = pool + p.offT[1];
    XYZZ<C> run = XYZZ<C>::inf(), acc = XYZZ<C>::inf();
    size_t base = (size_t)w * p.B + (size_t)s * p.L;
    bool any = false;
    for (int k = (int)p.L - 1; k >= 0; k--) {
        XYZZ<C> v = msm_bucket_value<C>(tstart, partials, base + k);
        if (!v.is_inf()) {
            run = padd_seg<C>(run, v);
            any = true;
        }
        if (any) acc = padd_seg<C>(acc, run);
    }
    segA[(size_t)w * p.S + s] = acc;
    segT[(size_t)w * p.S + s] = run;
}

// ---- level l >= 2: thread (w, g, role) ----------------------------------------------------------
// group g = items g*G .. g*G+G-1 of level l-1.
//   role 0:        T'_g = sum_j T_j ; new carried array C'_{l-1}[g] = sum_j j * T_j   (running sums)
//   role r, 1..l-1: C'_{r-1}[g] = sum_j C_{r-1}[j]
// With bucket index b = j_0 + L (j_1 + G_2 (j_2 + ... )):
//   sum_b (b+1) S_b = sum C_0 + L (sum C_1 + G_2 (sum C_2 + ... + G_K * sum_g g T_g))
template <class C>
HD void msm_level_body(const MsmPlan& p, int l, uint32_t w, uint32_t g, int role, XYZZ<C>* pool) {
    const uint32_t G = p.G[l], Sin = p.Sl[l - 1], Sout = p.Sl[l];
    const size_t in_per = (size_t)Sin * p.nwin, out_per = (size_t)Sout * p.nwin;
    const size_t ib = (size_t)w * Sin + (size_t)g * G, ob = (size_t)w * Sout + g;
    if (role > 0) {
        const XYZZ<C>* in = pool + p.offC[l - 1] + (size_t)(role - 1) * in_per;
        XYZZ<C> acc = in[ib];
        for (uint32_t j = 1; j < G; j++) acc = padd_call<C>(acc, in[ib + j]);
        pool[p.offC[l] + (size_t)(role - 1) * out_per + ob] = acc;
        return;
    }
    const XYZZ<C>* inT = pool + p.offT[l - 1];
    XYZZ<C> run = XYZZ<C>::inf(), acc = XYZZ<C>::inf();
    for (int j = (int)G - 1; j >= 1; j--) {
        run = padd_call<C>(run, inT[ib + j]);
        acc = padd_call<C>(acc, run);
    }
    pool[p.offC[l] + (size_t)(l - 1) * out_per + ob] = acc;
    pool[p.offT[l] + ob] = padd_call<C>(run, inT[ib]);
}

// ---- final: window value from the plane sums of the top level, then Horner over windows ---------
// bits[w * nplanes + r] = sum_g C_r[g] (r < nlev) ; bits[w * nplanes + nlev + k] = sum_{g : bit k of g} T_g
HD int msm_nplanes(const MsmPlan& p) { return p.nlev + p.logSK; }

template <class C>
HD XYZZ<C> msm_window_value(const MsmPlan& p, uint32_t w, const XYZZ<C>* bits) {
    const XYZZ<C>* bw = bits + (size_t)w * msm_nplanes(p);
    XYZZ<C> acc = XYZZ<C>::inf();
    for (int k = p.logSK - 1; k >= 0; k--) {
        acc = pdbl_call<C>(acc);
        acc = padd_call<C>(acc, bw[p.nlev + k]);
    }
    for (int l = p.nlev; l >= 1; l--) {
        for (uint32_t x = p.G[l]; x > 1; x >>= 1) acc = pdbl_call<C>(acc);
        acc = padd_call<C>(acc, bw[l - 1]);
    }
    return acc;
}

template <class C>
HD XYZZ<C> msm_horner_windows(const MsmPlan& p, const XYZZ<C>* R) {
    XYZZ<C> res = R[p.nwin - 1];
    for (int w = p.nwin - 2; w >= 0; w--) {
        if (!res.is_inf()) {
            if (C::A_ZERO) {
                Jac<C> j = to_jac<C>(res);
                for (int k = 0; k < p.c; k++) j = jdbl_a0<C>(j);
                res = from_jac<C>(j);
            } else {
                for (int k = 0; k < p.c; k++) res = pdbl_call<C>(res);
            }
        }
        res = padd_call<C>(res, R[w]);
    }
    return res;
}

}  // namespace ws
