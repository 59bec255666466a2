// host_math.cpp -- CPU-side unit-test shim: compiles the arithmetic headers' HOST path
// (portable C, see hd.cuh) so the group-law code can be checked against the oracle on a
// box without a GPU.  Test scaffolding only; the product library never takes this path.
#include <cstring>
#include "../../wasmsnark_b200/csrc/ec.cuh"
using namespace ws;

template <class T> static T ld(const uint8_t* p) { T t; memcpy(&t, p, sizeof(T)); return t; }
template <class T> static void st(uint8_t* p, const T& t) { memcpy(p, &t, sizeof(T)); }

extern "C" {
// op: 0 add, 1 sub, 2 mul, 3 inv(x), 4 neg(x), 5 to_mont, 6 from_mont ; field: 0 Fq, 1 Fr, 2 MNT6753 Fq, 3 BLS12-381 Fq, 4 BLS12-381 Fr
void h_fp_op(int field, int op, const uint8_t* x, const uint8_t* y, uint8_t* r) {
#define DO(P)                                                                    \
    {                                                                            \
        Fp<P> a = ld<Fp<P>>(x), b = ld<Fp<P>>(y), c;                             \
        switch (op) {                                                            \
            case 0: c = add(a, b); break;                                        \
            case 1: c = sub(a, b); break;                                        \
            case 2: c = mul(a, b); break;                                        \
            case 3: c = inv(a); break;                                           \
            case 4: c = neg(a); break;                                           \
            case 5: c = to_mont(a); break;                                       \
            default: c = from_mont(a); break;                                    \
        }                                                                        \
        st(r, c);                                                                \
    }
    if (field == 0) DO(BN128_Fq) else if (field == 1) DO(BN128_Fr) else if (field == 3) DO(BLS12_381_Fq) else if (field == 4) DO(BLS12_381_Fr) else DO(MNT6753_Fq)
#undef DO
}
void h_fq2_op(int op, const uint8_t* x, const uint8_t* y, uint8_t* r) {
    Fq2 a = ld<Fq2>(x), b = ld<Fq2>(y), c;
    switch (op) {
        case 0: c = add(a, b); break;
        case 1: c = sub(a, b); break;
        case 2: c = mul(a, b); break;
        case 3: c = inv(a); break;
        case 4: c = neg(a); break;
        default: c = sqr(a); break;
    }
    st(r, c);
}
// G1: acc (XYZZ 128 B) op affine (64 B) / XYZZ
void h_g1_madd(const uint8_t* acc, const uint8_t* q, uint8_t* r) { st(r, pmadd<G1Curve>(ld<G1X>(acc), ld<G1Affine>(q))); }
void h_g1_add(const uint8_t* a, const uint8_t* b, uint8_t* r) { st(r, padd<G1Curve>(ld<G1X>(a), ld<G1X>(b))); }
void h_g1_dbl(const uint8_t* a, uint8_t* r) { st(r, pdbl<G1Curve>(ld<G1X>(a))); }
void h_g1_from_affine(const uint8_t* q, uint8_t* r) { st(r, G1X::from_affine(ld<G1Affine>(q))); }
void h_g1_from_jac(const uint8_t* q, uint8_t* r) { st(r, from_jac<G1Curve>(ld<Jac<G1Curve>>(q))); }
void h_g1_to_affine(const uint8_t* a, uint8_t* r) { st(r, to_affine<G1Curve>(ld<G1X>(a))); }
void h_g1_times_scalar(const uint8_t* a, const uint8_t* k, int len, uint8_t* r) { st(r, times_scalar<G1Curve>(ld<G1X>(a), k, len)); }
void h_g2_madd(const uint8_t* acc, const uint8_t* q, uint8_t* r) { st(r, pmadd<G2Curve>(ld<G2X>(acc), ld<G2Affine>(q))); }
void h_g2_add(const uint8_t* a, const uint8_t* b, uint8_t* r) { st(r, padd<G2Curve>(ld<G2X>(a), ld<G2X>(b))); }
void h_g2_dbl(const uint8_t* a, uint8_t* r) { st(r, pdbl<G2Curve>(ld<G2X>(a))); }
void h_g2_from_affine(const uint8_t* q, uint8_t* r) { st(r, G2X::from_affine(ld<G2Affine>(q))); }
void h_g2_to_affine(const uint8_t* a, uint8_t* r) { st(r, to_affine<G2Curve>(ld<G2X>(a))); }
void h_g2_times_scalar(const uint8_t* a, const uint8_t* k, int len, uint8_t* r) { st(r, times_scalar<G2Curve>(ld<G2X>(a), k, len)); }
// BLS12-381: k * P + Q on affine Montgomery inputs (96 / 192 B), result as the affine Jacobian triple (x, y, 1) / (0, 1, 0)
void h_bls_g1_muladd(const uint8_t* p, const uint8_t* k, int len, const uint8_t* q, uint8_t* r) {
    XYZZ<BlsG1Curve> t = times_scalar<BlsG1Curve>(XYZZ<BlsG1Curve>::from_affine(ld<Affine<BlsG1Curve>>(p)), k, len);
    st(r, to_affine<BlsG1Curve>(pmadd<BlsG1Curve>(t, ld<Affine<BlsG1Curve>>(q))));
}
void h_bls_g2_muladd(const uint8_t* p, const uint8_t* k, int len, const uint8_t* q, uint8_t* r) {
    XYZZ<BlsG2Curve> t = times_scalar<BlsG2Curve>(XYZZ<BlsG2Curve>::from_affine(ld<Affine<BlsG2Curve>>(p)), k, len);
    st(r, to_affine<BlsG2Curve>(pmadd<BlsG2Curve>(t, ld<Affine<BlsG2Curve>>(q))));
}
}
