// api.cu -- C ABI (include/wasmsnark_b200.h) over the sm_100a kernels.  Plain CUDA runtime; no
// torch types cross this boundary.  There is no CPU implementation behind these entry points:
// without a device ws_create fails.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/wasmsnark_b200.h"
#include "ctx.cuh"
#include "field_kernels.cuh"

namespace ws {
void bls12381_launch_fp_batch(int field_is_fr, int op, const void* x, const void* y, void* out, size_t n, cudaStream_t st);   // bls12381_g1.cu
}
using namespace ws;

#define WS_API __attribute__((visibility("default")))

extern "C" {

WS_API const char* ws_version(void) { return "wasmsnark_b200 0.1 (sm_100a)"; }

WS_API int ws_create(int device, ws_ctx** out) {
    if (!out) return WS_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return WS_E_CUDA;
    ws_ctx* c = new ws_ctx();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return WS_E_CUDA;
    }
    for (int i = 0; i < StageTimer::MAX; i++) cudaEventCreate(&c->timer.ev[i]);
    // tuning overrides for A/B runs of whole programs (bench.py, tools/): same meaning as the ws_set_* hooks
    if (const char* e = getenv("WS_MSM_SEG_LEN")) {
        unsigned L = (unsigned)atoi(e);
        if (L && !(L & (L - 1))) c->seg_len = L;
    }
    if (const char* e = getenv("WS_MSM_ACC_VARIANT")) c->acc_variant = atoi(e);
    if (const char* e = getenv("WS_MSM_TASK_CAP")) c->force_T = (unsigned)atoi(e);
    if (const char* e = getenv("WS_NTT_VARIANT")) c->ntt_variant = atoi(e);
    *out = c;
    return WS_OK;
}

WS_API void ws_destroy(ws_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < StageTimer::MAX; i++) cudaEventDestroy(c->timer.ev[i]);
    if (c->cached_key && c->cached_key_free) c->cached_key_free(c->cached_key);
    for (auto& a : c->arenas)
        if (a.ptr) cudaFree(a.ptr);
    if (c->tw_table) cudaFree(c->tw_table);
    if (c->tw_table_bls) cudaFree(c->tw_table_bls);
    if (c->busy_ev) cudaEventDestroy(c->busy_ev);
    c->stager.destroy();
    if (c->lanes_ready) {
        for (int i = 0; i < ws_ctx::NLANES; i++) {
            cudaStreamDestroy(c->lanes[i]);
            cudaEventDestroy(c->lane_done[i]);
        }
        cudaEventDestroy(c->fork_ev);
        cudaEventDestroy(c->h_done);
    }
    cudaStreamDestroy(c->stream);
    ws::tls_error_forget(c);
    delete c;
}

// the calling thread's own last failure on this handle if it had one, else the handle's most recent failure
WS_API const char* ws_last_error(const ws_ctx* c) {
    if (!c) return "null ctx";
    const ws::TlsError& e = ws::tls_error();
    return e.owner == c ? e.text.c_str() : c->err.c_str();
}

WS_API unsigned long long ws_kernel_launches(const ws_ctx* c) { return c ? c->launches : 0; }

WS_API int ws_set_profiling(ws_ctx* c, int on) {
    WsLock ws_lock_(c);
    if (!c) return WS_E_ARG;
    c->timer.enabled = on != 0;
    return WS_OK;
}

WS_API int ws_last_timings(ws_ctx* c, const char** names, float* ms, int cap) {
    WsLock ws_lock_(c);
    if (!c) return WS_E_ARG;
    cudaSetDevice(c->device);
    if (cudaStreamSynchronize(c->timer.last_stream) != cudaSuccess) return c->fail_cuda("sync");
    int n = 0;
    for (int i = 1; i < c->timer.n && n < cap; i++) {
        float t = 0;
        cudaEventElapsedTime(&t, c->timer.ev[i - 1], c->timer.ev[i]);
        names[n] = c->timer.name[i];
        ms[n] = t;
        n++;
    }
    return n;
}

// ------------------------------------------------------------------ field batches
WS_API int ws_fp_batch(ws_ctx* c, int field, int op, const uint8_t* x, const uint8_t* y, uint8_t* out, size_t n) {
    WsLock ws_lock_(c);
    if (!c || !x || !out || (!y && (op == WS_OP_MUL || op == WS_OP_ADD || op == WS_OP_SUB))) return c ? c->fail(WS_E_ARG, "null buffer") : WS_E_ARG;
    if (field < 0 || field > WS_BLS12381_FR) return c->fail(WS_E_FIELD, "unknown field id");
    if (op < 0 || op > WS_OP_INVERSE) return c->fail(WS_E_ARG, "unknown op");
    if (n == 0) return WS_OK;
    cudaSetDevice(c->device);
    size_t esz = field == WS_MNT6753_FQ ? 96 : field == WS_BLS12381_FQ ? 48 : 32;
    uint8_t* d = (uint8_t*)c->arena(0, 3 * align256(n * esz));
    if (!d) return c->fail_cuda("workspace");
    uint8_t *dx = d, *dy = d + align256(n * esz), *dr = d + 2 * align256(n * esz);
    cudaMemcpyAsync(dx, x, n * esz, cudaMemcpyHostToDevice, c->stream);
    if (y) cudaMemcpyAsync(dy, y, n * esz, cudaMemcpyHostToDevice, c->stream);
    switch (field) {
        case WS_BN128_FQ: launch_fp_batch<BN128_Fq>(op, dx, dy, dr, n, c->stream); break;
        case WS_BN128_FR: launch_fp_batch<BN128_Fr>(op, dx, dy, dr, n, c->stream); break;
        case WS_BLS12381_FQ: bls12381_launch_fp_batch(0, op, dx, dy, dr, n, c->stream); break;
        case WS_BLS12381_FR: bls12381_launch_fp_batch(1, op, dx, dy, dr, n, c->stream); break;
        default: launch_fp_batch<MNT6753_Fq>(op, dx, dy, dr, n, c->stream); break;
    }
    cudaMemcpyAsync(out, dr, n * esz, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return c->fail_cuda("fp_batch");
    return WS_OK;
}

WS_API int ws_bn128_fq2_mul_batch(ws_ctx* c, const uint8_t* x, const uint8_t* y, uint8_t* out, size_t n) {
    WsLock ws_lock_(c);
    if (!c || !x || !y || !out) return c ? c->fail(WS_E_ARG, "null buffer") : WS_E_ARG;
    if (n == 0) return WS_OK;
    cudaSetDevice(c->device);
    uint8_t* d = (uint8_t*)c->arena(0, 3 * align256(n * 64));
    if (!d) return c->fail_cuda("workspace");
    uint8_t *dx = d, *dy = d + align256(n * 64), *dr = d + 2 * align256(n * 64);
    cudaMemcpyAsync(dx, x, n * 64, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(dy, y, n * 64, cudaMemcpyHostToDevice, c->stream);
    k_fq2_mul<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>((const Fq2*)dx, (const Fq2*)dy, (Fq2*)dr, n);
    cudaMemcpyAsync(out, dr, n * 64, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return c->fail_cuda("fq2_mul_batch");
    return WS_OK;
}

// test / tuning hook: force the task cap T (0 = automatic)
WS_API int ws_set_msm_task_cap(ws_ctx* c, unsigned T) {
    WsLock ws_lock_(c);
    if (!c) return WS_E_ARG;
    c->force_T = T;
    return WS_OK;
}

WS_API int ws_set_msm_acc_variant(ws_ctx* c, int v) {
    WsLock ws_lock_(c);
    if (!c) return WS_E_ARG;
    c->acc_variant = v;
    return WS_OK;
}

WS_API int ws_set_ntt_variant(ws_ctx* c, int v) {
    WsLock ws_lock_(c);
    if (!c) return WS_E_ARG;
    c->ntt_variant = v;
    return WS_OK;
}

// test / tuning hook: buckets per reduce segment (power of two)
WS_API int ws_set_msm_seg_len(ws_ctx* c, unsigned L) {
    WsLock ws_lock_(c);
    if (!c || (L & (L - 1))) return WS_E_ARG;   // 0 = automatic
    c->seg_len = L;
    return WS_OK;
}

}  // extern "C"
