// bls12381_g2.cu -- BLS12-381 G2 (over Fq2 = Fq[u]/(u^2+1), 2 x 12 limbs) multi-scalar multiplication.
//   g2m = a-zero Jacobian curve over f2m + g2m_multiexp   /root/reference/src/bls12381/build_bls12381.js:41-43
// Points: affine Montgomery x.c0|x.c1|y.c0|y.c1 (192 bytes), x == 0 => infinity; result Jacobian X|Y|Z (288 bytes).
// Parity: see bls12381_g1.cu ("parity unpinned": no reference image exists for this curve).
#include "msm_impl.cuh"
using namespace ws;

extern "C" {
WS_API int ws_bls12381_g2_multiexp(ws_ctx* c, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[288]) {
    WsLock ws_lock_(c);
    return multiexp_host<BlsG2Curve>(c, scalars, points, n, out);
}
WS_API int ws_bls12381_g2_multiexp_dev(ws_ctx* c, const void* ds, const void* dp, size_t n, void* dout, int wb, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !dout || (n && (!ds || !dp))) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    cudaSetDevice(c->device);
    return multiexp_dev<BlsG2Curve>(c, ds, dp, n, dout, wb, stream ? (cudaStream_t)stream : c->stream);
}
WS_API int ws_bls12381_g2_affine(ws_ctx* c, const uint8_t* in, uint8_t* out, size_t count, int from_montgomery) {
    WsLock ws_lock_(c);
    if (!c || !in || !out) return c ? c->fail(WS_E_ARG, "null buffer") : WS_E_ARG;
    if (count == 0) return WS_OK;
    cudaSetDevice(c->device);
    uint8_t* d = (uint8_t*)c->arena(0, 2 * align256(count * 288));
    if (!d) return c->fail_cuda("workspace");
    uint8_t* dout = d + align256(count * 288);
    cudaMemcpyAsync(d, in, count * 288, cudaMemcpyHostToDevice, c->stream);
    k_affine_points<BlsG2Curve><<<(unsigned)((count + 63) / 64), 64, 0, c->stream>>>((const Jac<BlsG2Curve>*)d, (Jac<BlsG2Curve>*)dout, (int)count, from_montgomery);
    cudaMemcpyAsync(out, dout, count * 288, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return c->fail_cuda("affine");
    return WS_OK;
}
}
