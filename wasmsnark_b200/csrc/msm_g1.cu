// msm_g1.cu -- BN128 G1 instantiation of the MSM pipeline + its C ABI entry points.
#include "msm_impl.cuh"
using namespace ws;

namespace ws {
int msm_g1_run(ws_ctx* c, const void* ds, const void* dp, size_t n, void* d_out_jac, void* d_out_xyzz, cudaStream_t st, int arena_slot, int table_c) {
    return multiexp_dev<G1Curve>(c, ds, dp, n, d_out_jac, 0, st, d_out_xyzz, false, arena_slot, table_c);
}
int msm_g1_build_table(ws_ctx* c, const void* d_points, size_t n, int table_c, void* d_table, cudaStream_t st) {
    return build_table_dev<G1Curve>(c, d_points, n, table_c, d_table, st);
}
int msm_table_windows_n(int table_c) { return msm_table_windows(table_c); }
int msm_table_choose_c(size_t n) {
    int cc;
    uint32_t T;
    msm_choose((uint32_t)n, &cc, &T, 256, 1);
    return cc;
}
}  // namespace ws

extern "C" {
WS_API int ws_bn128_g1_multiexp(ws_ctx* c, const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t out[96]) {
    WsLock ws_lock_(c);
    return multiexp_host<G1Curve>(c, scalars, points, n, out);
}
WS_API int ws_bn128_g1_multiexp_dev(ws_ctx* c, const void* ds, const void* dp, size_t n, void* dout, int wb, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !dout || (n && (!ds || !dp))) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    cudaSetDevice(c->device);
    return multiexp_dev<G1Curve>(c, ds, dp, n, dout, wb, stream ? (cudaStream_t)stream : c->stream);
}
WS_API int ws_bn128_g1_sum_dev(ws_ctx* c, const void* d_parts, int count, void* d_out, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !d_parts || !d_out || count < 0) return c ? c->fail(WS_E_ARG, "bad argument") : WS_E_ARG;
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    k_sum_points<G1Curve><<<1, 32, 0, st>>>((const Jac<G1Curve>*)d_parts, count, (Jac<G1Curve>*)d_out);
    return cudaGetLastError() == cudaSuccess ? WS_OK : c->fail_cuda("sum_points");
}
WS_API int ws_bn128_g1_affine(ws_ctx* c, const uint8_t* in, uint8_t* out, size_t count, int from_montgomery) {
    WsLock ws_lock_(c);
    if (!c || !in || !out) return c ? c->fail(WS_E_ARG, "null buffer") : WS_E_ARG;
    if (count == 0) return WS_OK;
    cudaSetDevice(c->device);
    uint8_t* d = (uint8_t*)c->arena(0, 2 * align256(count * 96));
    if (!d) return c->fail_cuda("workspace");
    uint8_t* dout = d + align256(count * 96);
    cudaMemcpyAsync(d, in, count * 96, cudaMemcpyHostToDevice, c->stream);
    k_affine_points<G1Curve><<<(unsigned)((count + 63) / 64), 64, 0, c->stream>>>((const Jac<G1Curve>*)d, (Jac<G1Curve>*)dout, (int)count, from_montgomery);
    cudaMemcpyAsync(out, dout, count * 96, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return c->fail_cuda("affine");
    return WS_OK;
}
WS_API int ws_bn128_g1_build_table_dev(ws_ctx* c, const void* d_points, size_t n, int table_c, void* d_table, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !d_points || !d_table) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    cudaSetDevice(c->device);
    return build_table_dev<G1Curve>(c, d_points, n, table_c, d_table, stream ? (cudaStream_t)stream : c->stream);
}
WS_API int ws_bn128_g1_multiexp_table_dev(ws_ctx* c, const void* ds, const void* d_table, size_t n, int table_c, void* dout, void* stream) {
    WsLock ws_lock_(c);
    if (!c || !dout || (n && (!ds || !d_table))) return c ? c->fail(WS_E_ARG, "null pointer") : WS_E_ARG;
    if (table_c < 8 || table_c > 22) return c->fail(WS_E_ARG, "table window bits must be in 8..22");
    cudaSetDevice(c->device);
    return multiexp_dev<G1Curve>(c, ds, d_table, n, dout, 0, stream ? (cudaStream_t)stream : c->stream, nullptr, true, 1, table_c);
}
WS_API int ws_msm_table_windows(int table_c) { return table_c >= 8 && table_c <= 22 ? msm_table_windows(table_c) : 0; }
}
