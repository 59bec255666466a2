// host_io.cuh -- getting caller-owned HOST buffers onto the device quickly.
//
// The drop-in call groth16GenProof(witness, provingKey) (src/bn128.js:580) hands the library ordinary, pageable
// memory (a Node ArrayBuffer, a Python bytes object).  cudaMemcpyAsync from pageable memory is staged by the
// driver on the calling thread at roughly one core's memcpy speed and blocks the caller; here the buffer is cut
// into slices, a few host threads copy their slices into a persistent pinned ring (two slots each) and queue the
// DMA of one slot while filling the other.  Pinned (cudaHostAlloc / cudaHostRegister) sources skip the staging.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

namespace ws {

struct PinnedStager {
    static constexpr int THREADS = 4;
    static constexpr size_t SLOT = (size_t)4 << 20;
    uint8_t* ring = nullptr;                    // THREADS x 2 slots
    cudaEvent_t ev[THREADS][2] = {};
    bool ready = false;

    bool init() {
        if (ready) return true;
        if (cudaHostAlloc((void**)&ring, THREADS * 2 * SLOT, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            ring = nullptr;
            return false;
        }
        for (int t = 0; t < THREADS; t++)
            for (int s = 0; s < 2; s++)
                if (cudaEventCreateWithFlags(&ev[t][s], cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) return false;
        ready = true;
        return true;
    }
    void destroy() {
        if (ring) cudaFreeHost(ring);
        ring = nullptr;
        if (ready)
            for (int t = 0; t < THREADS; t++)
                for (int s = 0; s < 2; s++) cudaEventDestroy(ev[t][s]);
        ready = false;
    }

    static bool is_pinned(const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeHost;
    }

    // Enqueues the copy of bytes from host `src` to device `dst` on `st` (device must be current).  Returns when the
    // source buffer has been read completely (the caller may reuse it); the device side is ordered on `st` as usual.
    cudaError_t upload(void* dst, const void* src, size_t bytes, cudaStream_t st, int device) {
        if (bytes == 0) return cudaSuccess;
        if (bytes < ((size_t)1 << 20) || is_pinned(src) || !init()) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
        const int nt = (int)std::min<size_t>(THREADS, (bytes + SLOT - 1) / SLOT);
        const size_t per = (((bytes + nt - 1) / nt) + 255) & ~(size_t)255;
        std::vector<cudaError_t> rc(nt, cudaSuccess);
        auto work = [&](int t) {
            cudaSetDevice(device);
            const size_t lo = std::min(bytes, per * t), hi = std::min(bytes, per * (t + 1));
            int slot = 0;
            bool used[2] = {false, false};
            for (size_t off = lo; off < hi; off += SLOT, slot ^= 1) {
                const size_t len = std::min(SLOT, hi - off);
                uint8_t* stage = ring + ((size_t)t * 2 + slot) * SLOT;
                if (used[slot]) cudaEventSynchronize(ev[t][slot]);          // the DMA that last read this slot is done
                memcpy(stage, (const uint8_t*)src + off, len);
                cudaError_t e = cudaMemcpyAsync((uint8_t*)dst + off, stage, len, cudaMemcpyHostToDevice, st);
                if (e != cudaSuccess) rc[t] = e;
                cudaEventRecord(ev[t][slot], st);
                used[slot] = true;
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        for (int t = 0; t < nt; t++)
            if (rc[t] != cudaSuccess) return rc[t];
        // the ring slots still feed DMAs in flight: the next upload's first touch of a slot waits on its event, but
        // the events must not be re-recorded before then -- they are per (thread, slot) and uploads on one context
        // are serialised by its mutex, so wait here for the tail DMAs (a few hundred microseconds at most)
        for (int t = 0; t < nt; t++)
            for (int s = 0; s < 2; s++) cudaEventSynchronize(ev[t][s]);
        return cudaSuccess;
    }
};

}  // namespace ws
