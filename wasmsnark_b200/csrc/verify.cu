// verify.cu -- host-side Groth16 verifier and pairing entry points of the C ABI (no GPU, no ws_ctx).
//
// Replaces Bn128.groth16Verify, /root/reference/src/bn128.js:722-786: vk_x = IC[0] + sum input_i * IC[i+1]
// (an input >= r rejects, :772), then e(A, B) * e(-vk_x, gamma2) * e(-C, delta2) * e(-alfa1, beta2) == 1 through
// bn128_pairingEq4 (:781).  All values cross the ABI as plain (non-Montgomery) little-endian integers -- the form
// verification_key.json / proof.json / public.json hold as decimal strings.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/wasmsnark_b200.h"
#include "pairing.cuh"

#define WS_API extern "C" __attribute__((visibility("default")))

namespace {
using namespace ws;
using namespace ws::pairing;

bool rd_fq(const uint8_t* b, Fq* out) {      // plain LE -> Montgomery; false if >= q
    Fq a;
    memcpy(a.v, b, 32);
    for (int i = 7; i >= 0; i--) {
        uint32_t m = BN128_Fq::mod(i);
        if (a.v[i] < m) break;
        if (a.v[i] > m || i == 0) return false;
    }
    *out = to_mont(a);
    return true;
}
void wr_fq(uint8_t* b, const Fq& a) {
    Fq p = from_mont(a);
    memcpy(b, p.v, 32);
}
bool rd_g1(const uint8_t* b, G1A* p) {
    if (!rd_fq(b, &p->x) || !rd_fq(b + 32, &p->y)) return false;
    p->inf = p->x.is_zero() && p->y.is_zero();
    return on_curve(*p);
}
bool rd_g2(const uint8_t* b, G2A* p) {
    if (!rd_fq(b, &p->x.c0) || !rd_fq(b + 32, &p->x.c1) || !rd_fq(b + 64, &p->y.c0) || !rd_fq(b + 96, &p->y.c1)) return false;
    p->inf = p->x.is_zero() && p->y.is_zero();
    return on_curve(*p);
}
G1A negate(const G1A& p) { return G1A{p.x, neg(p.y), p.inf}; }
}  // namespace

WS_API int ws_bn128_pairing(const uint8_t p[64], const uint8_t q[128], uint8_t out[384]) {
    if (!p || !q || !out) return WS_E_ARG;
    G1A P;
    G2A Q;
    if (!rd_g1(p, &P) || !rd_g2(q, &Q)) return WS_E_ARG;
    bool ok = true;
    F12 e = ws::pairing::pairing(P, Q, &ok);
    if (!ok) return WS_E_ARG;          // Q outside the order-r subgroup of the twist (vertical line in the Miller loop)
    for (int i = 0; i < 6; i++) {   // reference ftm layout: c0.(a0,a1,a2), c1.(a0,a1,a2), each (re, im)
        const F6& h = i < 3 ? e.c0 : e.c1;
        const F2& a = (i % 3) == 0 ? h.a0 : (i % 3) == 1 ? h.a1 : h.a2;
        wr_fq(out + 64 * i, a.c0);
        wr_fq(out + 64 * i + 32, a.c1);
    }
    return WS_OK;
}

WS_API int ws_bn128_groth16_verify(const uint8_t vk_alfa_1[64], const uint8_t vk_beta_2[128], const uint8_t vk_gamma_2[128],
                                   const uint8_t vk_delta_2[128], const uint8_t* ic, size_t n_ic, const uint8_t* inputs, size_t n_inputs,
                                   const uint8_t pi_a[64], const uint8_t pi_b[128], const uint8_t pi_c[64], int* valid) {
    if (!vk_alfa_1 || !vk_beta_2 || !vk_gamma_2 || !vk_delta_2 || !ic || !pi_a || !pi_b || !pi_c || !valid || (n_inputs && !inputs)) return WS_E_ARG;
    if (n_ic != n_inputs + 1) return WS_E_ARG;
    *valid = 0;
    G1A A, C, alfa, icp;
    G2A B, beta, gamma, delta;
    // malformed points (coordinate >= q, not on the curve) can never satisfy the equation: verdict false
    if (!rd_g1(pi_a, &A) || !rd_g2(pi_b, &B) || !rd_g1(pi_c, &C) || !rd_g1(vk_alfa_1, &alfa) || !rd_g2(vk_beta_2, &beta) ||
        !rd_g2(vk_gamma_2, &gamma) || !rd_g2(vk_delta_2, &delta))
        return WS_OK;
    if (!rd_g1(ic, &icp)) return WS_OK;
    XYZZ<G1Curve> acc = icp.inf ? XYZZ<G1Curve>::inf() : XYZZ<G1Curve>{icp.x, icp.y, Fq::one(), Fq::one()};
    std::vector<Affine<G1Curve>> pts;
    std::vector<const uint8_t*> scs;
    for (size_t i = 0; i < n_inputs; i++) {
        const uint8_t* s = inputs + 32 * i;
        uint32_t w[8];
        memcpy(w, s, 32);
        bool lt = false;                              // input must be < r  (src/bn128.js:772)
        for (int k = 7; k >= 0; k--) {
            uint32_t m = BN128_Fr::mod(k);
            if (w[k] != m) {
                lt = w[k] < m;
                break;
            }
        }
        if (!lt) return WS_OK;
        if (!rd_g1(ic + 64 * (i + 1), &icp)) return WS_OK;
        if (icp.inf) continue;
        pts.push_back(Affine<G1Curve>{icp.x, icp.y});
        scs.push_back(s);
    }
    // vk_x = IC[0] + sum input_i IC[i+1].  The reference adds g1m_timesScalar results one by one (:774-776); the same
    // group element comes out of one small bucket pass: 4-bit windows, 15 buckets, mixed additions.
    {
        constexpr int C = 4, NB = (1 << C) - 1;
        XYZZ<G1Curve> sum = XYZZ<G1Curve>::inf();
        for (int win = 256 / C - 1; win >= 0; win--) {
            for (int d = 0; d < C; d++) sum = pdbl_call<G1Curve>(sum);
            XYZZ<G1Curve> bucket[NB];
            bool used = false;
            for (int b = 0; b < NB; b++) bucket[b] = XYZZ<G1Curve>::inf();
            for (size_t i = 0; i < pts.size(); i++) {
                unsigned dig = (scs[i][win >> 1] >> ((win & 1) * 4)) & 15u;
                if (!dig) continue;
                bucket[dig - 1] = pmadd<G1Curve>(bucket[dig - 1], pts[i]);
                used = true;
            }
            if (!used) continue;
            XYZZ<G1Curve> run = XYZZ<G1Curve>::inf(), tot = XYZZ<G1Curve>::inf();
            for (int b = NB - 1; b >= 0; b--) {       // sum_b (b + 1) bucket[b] by running sums
                run = padd_call<G1Curve>(run, bucket[b]);
                tot = padd_call<G1Curve>(tot, run);
            }
            sum = padd_call<G1Curve>(sum, tot);
        }
        acc = padd_call<G1Curve>(acc, sum);
    }
    Jac<G1Curve> ax = to_affine<G1Curve>(acc);
    G1A vkx{ax.x, ax.y, ax.z.is_zero()};
    G1A Ps[4] = {A, negate(vkx), negate(C), negate(alfa)};
    G2A Qs[4] = {B, gamma, delta, beta};
    bool ok = true;
    F12 ml = miller_product(Ps, Qs, 4, &ok);
    if (!ok) return WS_OK;             // a G2 input outside the order-r subgroup: reject (*valid stays 0)
    *valid = (final_exponentiation(ml) == F12::one()) ? 1 : 0;
    return WS_OK;
}
