// ctx.cuh -- the library context: device, stream, growable HBM arenas, stage timer.
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <string>

#include "../../include/wasmsnark_b200.h"
#include "host_io.cuh"
#include "tls_error.cuh"

namespace ws {

// CUDA-event stamps between pipeline stages, on the stream the kernels are launched on.
struct StageTimer {
    static constexpr int MAX = 48;
    cudaEvent_t ev[MAX];
    const char* name[MAX];
    int n = 0;
    bool enabled = false;
    cudaStream_t last_stream = nullptr;
    void begin(cudaStream_t st) {
        n = 0;
        last_stream = st;
        mark("begin", st);
    }
    void mark(const char* nm, cudaStream_t st) {
        if (!enabled || n >= MAX) return;
        name[n] = nm;
        cudaEventRecord(ev[n], st);
        n++;
    }
};

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct Arena {
    void* ptr = nullptr;
    size_t size = 0;
};

}  // namespace ws

struct ws_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // Thread safety (the reference queues any number of commands, src/bn128.js:282-300): every C-ABI entry point
    // holds `mu` for the duration of the call, so host threads sharing a context are serialised -- blocking calls
    // run one after the other; asynchronous *_dev calls are additionally chained on the device through `busy_ev`
    // (they share the context's workspace arenas): work enqueued on a different stream than the previous call's
    // first waits for that call's work.
    std::recursive_mutex mu;
    cudaEvent_t busy_ev = nullptr;
    cudaStream_t busy_stream = nullptr;
    bool busy = false;
    void enter_stream(cudaStream_t st) {
        if (busy && busy_stream != st) cudaStreamWaitEvent(st, busy_ev, 0);
    }
    void leave_stream(cudaStream_t st) {
        if (!busy_ev && cudaEventCreateWithFlags(&busy_ev, cudaEventDisableTiming) != cudaSuccess) {
            busy_ev = nullptr;
            return;
        }
        cudaEventRecord(busy_ev, st);
        busy_stream = st;
        busy = true;
    }
    ws::Arena arenas[16];
    // side streams for the prover: independent MSMs and the r/s fixed-base products run concurrently
    static constexpr int NLANES = 6;
    cudaStream_t lanes[NLANES] = {};
    cudaEvent_t lane_done[NLANES] = {};
    cudaEvent_t fork_ev = nullptr;
    bool lanes_ready = false;
    bool pre_pending = false;   // prover: (r, s) products already in flight on lanes[4]
    bool h_priority = false;    // prover: run the H chain on the high-priority lane (ws_set_h_priority)
    cudaEvent_t h_done = nullptr;
    bool ensure_lanes() {
        if (lanes_ready) return true;
        int prio_least = 0, prio_greatest = 0;
        cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
        for (int i = 0; i < NLANES; i++) {
            // lanes[5] carries the H chain (LC, five NTTs, H MSM) at the highest priority: its blocks are scheduled
            // ahead of the four witness MSMs on lanes 0..3, so the one dependent chain of a proof is not stretched
            if (cudaStreamCreateWithPriority(&lanes[i], cudaStreamNonBlocking, i == 5 ? prio_greatest : prio_least) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&lane_done[i], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        if (cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&h_done, cudaEventDisableTiming) != cudaSuccess) return false;
        lanes_ready = true;
        return true;
    }
    ws::StageTimer timer;
    ws::PinnedStager stager;     // pageable caller buffers -> device (host_io.cuh)
    std::string err;
    unsigned long long launches = 0;   // kernels of this library launched so far (ws_kernel_launches)
    unsigned force_T = 0;
    // one-entry proving-key cache of the stateless prove call, keyed by a content fingerprint (prover.cu)
    void* cached_key = nullptr;
    unsigned long long cached_fp = 0;
    size_t cached_len = 0;
    void (*cached_key_free)(void*) = nullptr;
    bool key_cache = true;
    bool key_tables = true;      // resident proving keys get window tables (prover.cu)
    void* tw_table = nullptr;     // ntt.cuh TwiddleTable storage (Fr*), owned here
    void* tw_table_bls = nullptr; // same for BLS12-381 Fr (bls12381_g1.cu)
    int tw_logM_bls = 0;
    int tw_logM = 0;
    int acc_variant = 0;         // tuning hook: launch-bounds variant of the accumulate kernel
    int ntt_variant = 0;         // tuning hook: occupancy variant of the radix-64 NTT pass (ws_set_ntt_variant)
    unsigned seg_len = 0;    // 0 = automatic

    // Arena `slot` grown to at least `bytes` (contents are scratch: not preserved on growth).
    void* arena(int slot, size_t bytes) {
        ws::Arena& a = arenas[slot];
        if (a.size >= bytes && a.ptr) return a.ptr;
        if (a.ptr) {
            cudaDeviceSynchronize();      // work on any stream may still use the old block
            cudaFree(a.ptr);
            a.ptr = nullptr;
            a.size = 0;
        }
        size_t want = bytes + bytes / 8 + 4096;
        if (cudaMalloc(&a.ptr, want) != cudaSuccess) {
            a.ptr = nullptr;
            return nullptr;
        }
        a.size = want;
        return a.ptr;
    }
    int fail(int code, const char* msg) {
        err = msg;
        ws::tls_error_set(this, err);
        return code;
    }
    int fail_cuda(const char* where) {
        cudaError_t e = cudaGetLastError();
        err = std::string(where) + ": " + cudaGetErrorString(e);
        ws::tls_error_set(this, err);
        return WS_E_CUDA;
    }
};

// Holds the context's mutex for the duration of one C-ABI call (null ctx: no-op, the caller returns WS_E_ARG).
struct WsLock {
    ws_ctx* c;
    explicit WsLock(ws_ctx* c_) : c(c_) {
        if (c) c->mu.lock();
    }
    ~WsLock() {
        if (c) c->mu.unlock();
    }
    WsLock(const WsLock&) = delete;
    WsLock& operator=(const WsLock&) = delete;
};
// Orders one asynchronous call's work after the previous call's on the device (see ws_ctx::busy_ev).
struct WsStreamScope {
    ws_ctx* c;
    cudaStream_t st;
    WsStreamScope(ws_ctx* c_, cudaStream_t st_) : c(c_), st(st_) { c->enter_stream(st); }
    ~WsStreamScope() { c->leave_stream(st); }
    WsStreamScope(const WsStreamScope&) = delete;
    WsStreamScope& operator=(const WsStreamScope&) = delete;
};
