// tests/host/cuemu/cuda_runtime.h -- TEST SCAFFOLDING: a stand-in for <cuda_runtime.h> that lets the library's .cu sources be
// compiled by plain g++ and their kernels be EXECUTED ON THE CPU, thread by thread (tests/host/cuemu/README in cuemu.cpp).
//
// Why: the build container has no GPU, so everything above the per-thread arithmetic -- launch geometry, shared-memory
// exchanges, barriers, warp shuffles, scans, the stream / event choreography of the prover and of the multi-GPU handle --
// could only be exercised on the GPU box.  With this header the SAME kernel source runs as one fiber per CUDA thread
// (blocks one after the other, threads of a block round-robin, __syncthreads / __shfl_*_sync / __syncwarp as rendezvous
// points), "device memory" is host memory, and the runtime calls are synchronous.  The field arithmetic takes its portable
// host path (fp.cuh: no __CUDA_ARCH__), so the PTX carry chains are NOT covered here (tests/test_fp_asm_model.py and the GPU
// tests cover those).  The result, tests/host/libwasmsnark_b200_emu.so, is loaded only by tests
// (tests/test_emulated_library.py); the package never looks for it -- it is not a CPU fallback.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <tuple>

#define __CUDACC__ 1          // the sources' "compiled as CUDA" paths (hd.cuh, warp_ec.cuh) ...
#undef __CUDA_ARCH__          // ... but never their device-only PTX paths

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
// __noinline__ is rewritten in the copied sources (build_emu.py): a macro of that name would break __attribute__((__noinline__)) in system headers
#define __launch_bounds__(...)
#define __shared__ static thread_local          // one copy per host thread: every block of a launch runs on the launching thread

struct uint3 { unsigned x, y, z; };
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace cuemu {
extern thread_local uint3 tl_threadIdx, tl_blockIdx;
extern thread_local dim3 tl_blockDim, tl_gridDim;
void launch(dim3 grid, dim3 block, size_t dyn_smem, const char* kernel, const std::function<void()>& body);     // runs the whole grid, returns when done
void count_op(int limbs, int squaring);       // operation census (cuemu.cpp)
// "k_msm_digits<C>" + "... [with C = ws::G1Curve; ...]" -> "k_msm_digits<C> [C = G1Curve]"
static inline std::string kernel_label(const char* kernel, const char* site) {
    std::string s(kernel), w(site);
    size_t i = w.find("[with ");
    if (i == std::string::npos) return s;
    std::string b = w.substr(i + 6);
    if (!b.empty() && b.back() == ']') b.pop_back();
    for (size_t k; (k = b.find("ws::")) != std::string::npos;) b.erase(k, 4);
    // keep only the type bindings (drop "cudaStream_t = ..." and similar aliases)
    std::string keep;
    size_t pos = 0;
    while (pos < b.size()) {
        size_t e = b.find("; ", pos);
        std::string item = b.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
        if (item.find("cudaStream_t") == std::string::npos && item.find("typename") == std::string::npos) keep += (keep.empty() ? "" : "; ") + item;
        if (e == std::string::npos) break;
        pos = e + 2;
    }
    return keep.empty() ? s : s + " [" + keep + "]";
}
void* dyn_smem();                 // the current block's dynamic shared memory
void sync_block();                // __syncthreads
void sync_warp();                 // __syncwarp
uint32_t shfl(uint32_t v, int src_lane);      // value of lane src_lane (of the caller's warp); a rendezvous of the whole warp
int lane();
}  // namespace cuemu

#define threadIdx (cuemu::tl_threadIdx)
#define blockIdx (cuemu::tl_blockIdx)
#define blockDim (cuemu::tl_blockDim)
#define gridDim (cuemu::tl_gridDim)
static const int warpSize = 32;

static inline void __syncthreads() { cuemu::sync_block(); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { cuemu::sync_warp(); }
// all uses in the library pass the full mask and width 32; lanes outside the warp's live set read their own value
static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) { return cuemu::shfl(v, src & 31); }
static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned d) {
    int l = cuemu::lane();
    uint32_t r = cuemu::shfl(v, l >= (int)d ? l - (int)d : l);
    return l >= (int)d ? r : v;
}
static inline uint32_t __shfl_down_sync(unsigned, uint32_t v, unsigned d) {
    int l = cuemu::lane();
    uint32_t r = cuemu::shfl(v, l + (int)d < 32 ? l + (int)d : l);
    return l + (int)d < 32 ? r : v;
}
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int m) { return cuemu::shfl(v, (cuemu::lane() ^ m) & 31); }
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
    s &= 31;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
// blocks run one after the other and the threads of a block interleave only at rendezvous points: plain read-modify-write
template <class T>
static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T>
static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T>
static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T>
static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T>
static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
using std::max;
using std::min;

// ------------------------------------------------------------------------------------------- runtime API (synchronous)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorInvalidDevice = 101 };
typedef struct cuemu_stream* cudaStream_t;
typedef struct cuemu_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamDefault = 0, cudaStreamNonBlocking = 1 };
enum { cudaEventDefault = 0, cudaEventBlockingSync = 1, cudaEventDisableTiming = 2 };
enum { cudaHostAllocDefault = 0, cudaHostRegisterDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes {
    cudaMemoryType type;
    int device;
    void* devicePointer;
    void* hostPointer;
};

extern "C" {
cudaError_t cudaGetDeviceCount(int* n);            // CUEMU_DEVICES (default 1) pretend devices
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaMalloc(void** p, size_t n);
cudaError_t cudaFree(void* p);
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned flags);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaHostRegister(void* p, size_t n, unsigned flags);
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaMemcpyPeerAsync(void* d, int dd, const void* s, int sd, size_t n, cudaStream_t st = nullptr);
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st = nullptr);
cudaError_t cudaMemset(void* d, int v, size_t n);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned flags, int prio);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest);
cudaError_t cudaDeviceCanAccessPeer(int* can, int d, int p);
cudaError_t cudaDeviceEnablePeerAccess(int p, unsigned flags);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetLastError(void);
cudaError_t cudaPeekAtLastError(void);
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p);
}
template <class T>
static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }

namespace cuemu {
// a launch on a stream: `body` owns copies of the kernel's arguments, evaluated at the launch (build_emu.py).  Runs at once
// in the synchronous mode, is queued on `st` under CUEMU_ASYNC (cuemu.cpp)
void launch_on(cudaStream_t st, dim3 grid, dim3 block, size_t dyn_smem, const std::string& kernel, std::function<void()> body);
}  // namespace cuemu
