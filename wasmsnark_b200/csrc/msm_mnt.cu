// msm_mnt.cu -- MNT6753 (768-bit, 24-limb) wide-limb path: Fq3 products and the G1 multi-scalar
// multiplication of BASELINE config C5.
//
// The reference has the field / extension / curve modules for this curve
//   f1m_* over q (n64 = 12)     /root/reference/src/mnt6753/build_mnt6753.js:19-31  (src/build_f1m.js)
//   f3m_mul, u^3 = 11           /root/reference/src/build_f3m.js:125-189, build_mnt6753.js:35-56
//   g1m_* projective, a = 11    /root/reference/src/build_curve_projective.js, build_mnt6753.js:90-109
// but NO multiexp for it (build_mnt6753.js never calls buildMultiexp).  The MSM here is therefore
// checked against the definitional sum  sum_i g1m_timesScalar(P_i, s_i)  folded with g1m_add,
// computed by the reference image's own functions -- "parity vs definitional sum, reference MSM absent".
// Input conventions follow the BN128 path: raw little-endian scalars (96 bytes), affine Montgomery
// points x|y (192 bytes), x == 0 meaning infinity; result = the reference's projective triple X|Y|Z.
#include "msm_impl.cuh"
using namespace ws;

namespace ws {
struct MntFq3 {
    MntFq c0, c1, c2;
};
}  // namespace ws

static __global__ void __launch_bounds__(64) k_mnt_fq3_mul(const MntFq3* __restrict__ x, const MntFq3* __restrict__ y, MntFq3* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    MntFq3 a = ldg_struct(x + i), b = ldg_struct(y + i), r;
    // build_f3m.js:158-186 with mulNonResidue = multiplication by 11
    MntFq aA = mul(a.c0, b.c0), bB = mul(a.c1, b.c1), cC = mul(a.c2, b.c2);
    MntFq t0 = sub(mul(add(a.c1, a.c2), add(b.c1, b.c2)), add(bB, cC));
    r.c0 = add(aA, MntG1Curve::mul_a(t0));
    MntFq t1 = sub(mul(add(a.c0, a.c1), add(b.c0, b.c1)), add(aA, bB));
    r.c1 = add(t1, MntG1Curve::mul_a(cC));
    MntFq t2 = sub(mul(add(a.c0, a.c2), add(b.c0, b.c2)), add(aA, cC));
    r.c2 = add(t2, bB);
    stg_struct(out + i, r);
}

extern "C" {
WS_API int ws_mnt6753_fq3_mul_batch(ws_ctx* c, const uint8_t* x, const uint8_t* y, uint8_t* out, size_t n) {
    WsLock ws_lock_(c);
    if (!c || !x || !y || !out) return c ? c->fail(WS_E_ARG, "null buffer") : WS_E_ARG;
    if (n == 0) return WS_OK;
    cudaSetDevice(c->device);
    uint8_t* d = (uint8_t*)c->arena(0, 3 * align256(n * 288));
    if (!d) return c->fail_cuda("workspace");
    uint8_t *dx = d, *dy = d + align256(n * 288), *dr = d + 2 * align256(n * 288);
    cudaMemcpyAsync(dx, x, n * 288, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(dy, y, n * 288, cudaMemcpyHostToDevice, c->stream);
    k_mnt_fq3_mul<<<(unsigned)((n + 63) / 64), 64, 0, c->stream>>>((const MntFq3*)dx, (const MntFq3*)dy, (MntFq3*)dr, n);
    cudaMemcpyAsync(out, dr, n * 288, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return c->fail_cuda("fq3_mul_batch");
    return WS_OK;
}
WS_API int ws_mnt6753_g1_multiexp(ws_ctx* c, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[288]) {
    WsLock ws_lock_(c);
    return multiexp_host<MntG1Curve>(c, scalars, points, n, out);
}
WS_API int ws_mnt6753_g1_multiexp_dev(ws_ctx* c, const void* ds, const void* dp, size_t n, void* dout, int wb, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !dout || (n && (!ds || !dp))) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    cudaSetDevice(c->device);
    return multiexp_dev<MntG1Curve>(c, ds, dp, n, dout, wb, stream ? (cudaStream_t)stream : c->stream);
}
WS_API int ws_mnt6753_g1_affine(ws_ctx* c, const uint8_t* in, uint8_t* out, size_t count, int from_montgomery) {
    WsLock ws_lock_(c);
    if (!c || !in || !out) return c ? c->fail(WS_E_ARG, "null buffer") : WS_E_ARG;
    if (count == 0) return WS_OK;
    cudaSetDevice(c->device);
    uint8_t* d = (uint8_t*)c->arena(0, 2 * align256(count * 288));
    if (!d) return c->fail_cuda("workspace");
    uint8_t* dout = d + align256(count * 288);
    cudaMemcpyAsync(d, in, count * 288, cudaMemcpyHostToDevice, c->stream);
    k_affine_points<MntG1Curve><<<(unsigned)((count + 63) / 64), 64, 0, c->stream>>>((const Jac<MntG1Curve>*)d, (Jac<MntG1Curve>*)dout, (int)count, from_montgomery);
    cudaMemcpyAsync(out, dout, count * 288, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return c->fail_cuda("affine");
    return WS_OK;
}
}
