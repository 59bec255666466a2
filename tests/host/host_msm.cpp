// host_msm.cpp -- sequential CPU emulation of the MSM pipeline's index logic
// (msm_core.cuh bodies run thread by thread; the cooperative steps -- scan, warp combine,
// block bit-sums -- are replaced by their serial definitions).  Test scaffolding only.
#include <cstring>
#include <vector>
#include "../../wasmsnark_b200/csrc/msm_core.cuh"
using namespace ws;

template <class C>
static void run(const uint8_t* scalars, const uint8_t* points, uint32_t n, int c, uint32_t T, uint8_t* out_affine, int table = 0) {
    MsmPlan p = msm_make_plan(n, c, T, 0, 256, table);
    const Affine<C>* pts = (const Affine<C>*)points;
    std::vector<Affine<C>> tab;
    if (table) {      // window table T[w * n + i] = 2^(c w) * P_i, affine, x = 0 for infinity (what msm_build_table makes on the device)
        tab.resize((size_t)p.ndig * n);
        for (uint32_t i = 0; i < n; i++) {
            XYZZ<C> cur = XYZZ<C>::from_affine(pts[i]);
            for (int w = 0; w < p.ndig; w++) {
                Jac<C> a = to_affine<C>(cur);
                Affine<C> e{a.x, a.y};
                if (cur.is_inf()) e = Affine<C>{C::F::zero(), C::F::zero()};
                tab[(size_t)w * n + i] = e;
                for (int k = 0; k < c; k++) cur = pdbl<C>(cur);
            }
        }
    }
    const Affine<C>* gather = table ? tab.data() : pts;
    std::vector<int32_t> digits((size_t)p.ndig * n);
    std::vector<uint32_t> counts(p.nbuckets, 0), bstart(p.nbuckets + 1), tstart(p.nbuckets + 1);
    for (uint32_t i = 0; i < n; i++) msm_digits_body<8>(p, i, (const uint32_t*)scalars + (size_t)i * 8, pts[i].is_inf(), digits.data(), counts.data());
    bstart[0] = 0; tstart[0] = 0;
    for (uint32_t b = 0; b < p.nbuckets; b++) {
        bstart[b + 1] = bstart[b] + counts[b];
        tstart[b + 1] = tstart[b] + msm_ntasks(p, counts[b]);
    }
    std::vector<uint32_t> entries(bstart[p.nbuckets] + 1), task_bucket(p.max_tasks), heavy(2 * (size_t)p.nbuckets), len_hist(p.T + 2, 0), len_cursor(p.T + 2, 0), order(p.max_tasks);
    uint32_t nheavy[2] = {0, 0};
    for (uint32_t i = 0; i < n; i++) msm_scatter_body(p, i, digits.data(), bstart.data(), counts.data(), entries.data());
    for (uint32_t b = 0; b < p.nbuckets; b++) msm_filltasks_body(p, b, bstart.data(), tstart.data(), task_bucket.data(), len_hist.data(), heavy.data(), nheavy);
    uint32_t ntasks = tstart[p.nbuckets];
    msm_task_offsets_body(p, len_hist.data(), len_cursor.data());
    for (uint32_t t = 0; t < ntasks; t++) msm_task_order_body(p, t, bstart.data(), tstart.data(), task_bucket.data(), len_cursor.data(), order.data());
    std::vector<XYZZ<C>> partials(ntasks + 1);
    for (uint32_t i = 0; i < ntasks; i++) partials[order[i]] = msm_accumulate_body<C>(p, order[i], bstart.data(), tstart.data(), task_bucket.data(), entries.data(), gather);
    for (uint32_t h = 0; h < nheavy[0] + nheavy[1]; h++) {   // combine (serial stand-in for the warp / block trees)
        uint32_t b = h < nheavy[0] ? heavy[h] : heavy[p.nbuckets + (h - nheavy[0])];
        XYZZ<C> acc = partials[tstart[b]];
        for (uint32_t t = tstart[b] + 1; t < tstart[b + 1]; t++) acc = padd<C>(acc, partials[t]);
        partials[tstart[b]] = acc;
    }
    std::vector<XYZZ<C>> pool(p.pool_elems + 1), bits((size_t)p.nwin * msm_nplanes(p)), R(p.nwin);
    for (int w = 0; w < p.nwin; w++)
        for (uint32_t s = 0; s < p.S; s++) msm_segreduce_body<C>(p, w, s, tstart.data(), partials.data(), pool.data());
    for (int l = 2; l <= p.nlev; l++)
        for (int w = 0; w < p.nwin; w++)
            for (uint32_t g = 0; g < p.Sl[l]; g++)
                for (int role = 0; role < l; role++) msm_level_body<C>(p, l, w, g, role, pool.data());
    for (int w = 0; w < p.nwin; w++) {        // plane sums over the top level (serial stand-in for the warp reduction)
        XYZZ<C>* bw = &bits[(size_t)w * msm_nplanes(p)];
        for (int k = 0; k < msm_nplanes(p); k++) bw[k] = XYZZ<C>::inf();
        const uint32_t SK = p.Sl[p.nlev];
        const size_t per = (size_t)SK * p.nwin;
        for (uint32_t g = 0; g < SK; g++) {
            size_t o = (size_t)w * SK + g;
            for (int r = 0; r < p.nlev; r++) bw[r] = padd<C>(bw[r], pool[p.offC[p.nlev] + (size_t)r * per + o]);
            for (int k = 0; k < p.logSK; k++)
                if ((g >> k) & 1) bw[p.nlev + k] = padd<C>(bw[p.nlev + k], pool[p.offT[p.nlev] + o]);
        }
    }
    for (int w = 0; w < p.nwin; w++) R[w] = msm_window_value<C>(p, w, bits.data());
    XYZZ<C> res = msm_horner_windows<C>(p, R.data());
    Jac<C> a = to_affine<C>(res);
    memcpy(out_affine, &a, sizeof(a));
}

extern "C" {
void h_msm_g1(const uint8_t* scalars, const uint8_t* points, uint32_t n, int c, uint32_t T, uint8_t* out) { run<G1Curve>(scalars, points, n, c, T, out); }
void h_msm_g2(const uint8_t* scalars, const uint8_t* points, uint32_t n, int c, uint32_t T, uint8_t* out) { run<G2Curve>(scalars, points, n, c, T, out); }
void h_msm_g1_table(const uint8_t* scalars, const uint8_t* points, uint32_t n, int c, uint32_t T, uint8_t* out) { run<G1Curve>(scalars, points, n, c, T, out, 1); }
void h_msm_g2_table(const uint8_t* scalars, const uint8_t* points, uint32_t n, int c, uint32_t T, uint8_t* out) { run<G2Curve>(scalars, points, n, c, T, out, 1); }
}
