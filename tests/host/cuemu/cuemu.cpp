// tests/host/cuemu/cuemu.cpp -- TEST SCAFFOLDING: executes CUDA kernels on the CPU (see cuda_runtime.h in this directory).
//
// Execution model.  launch(grid, block, smem, body) runs on the calling host thread and returns when the grid is done:
// blocks run one after the other; inside a block every CUDA thread is a fiber (own stack, x86-64 context switch below),
// scheduled round-robin; a fiber gives up the CPU only at a rendezvous -- __syncthreads (all live threads of the block),
// __syncwarp / __shfl_*_sync (all live lanes of its warp).  A thread that returned no longer takes part in rendezvous
// (what the hardware does for the kernels of this library: early-exit guards sit in front of the barriers only where whole
// warps / blocks leave together).  A pass over all fibers without any progress is a deadlock and aborts with a message.
// Launches from several host threads (the multi-GPU handle, the staging threads) are independent: all scheduler state is
// thread_local.  "Device memory" is host memory; the runtime API below is synchronous and never fails except on bad arguments.
#include "cuda_runtime.h"

#include <sys/mman.h>

#include <atomic>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#if !defined(__x86_64__)
#error "cuemu's context switch is written for x86-64"
#endif

// void cuemu_switch(void** save_sp, void* load_sp): saves the callee-saved registers and the stack pointer of the running
// context, continues the other one
asm(R"(
.text
.globl cuemu_switch
.type cuemu_switch,@function
cuemu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cuemu_switch,.-cuemu_switch
)");
extern "C" void cuemu_switch(void** save_sp, void* load_sp);

// AddressSanitizer build (build_emu.py --asan: out-of-bounds accesses of kernels and host code become reports): the
// sanitizer has to be told about every stack switch
#ifdef CUEMU_ASAN
extern "C" void __sanitizer_start_switch_fiber(void** fake_stack_save, const void* bottom, size_t size);
extern "C" void __sanitizer_finish_switch_fiber(void* fake_stack_save, const void** bottom_old, size_t* size_old);
extern "C" void __asan_unpoison_memory_region(void const volatile* addr, size_t size);
#endif

namespace cuemu {

thread_local int tl_last_error = 0;      // sticky until read, like the runtime's
thread_local uint3 tl_threadIdx, tl_blockIdx;
thread_local dim3 tl_blockDim, tl_gridDim;

namespace {
constexpr size_t STACK_BYTES = 256 << 10;

struct Fiber {
    void* sp = nullptr;
    void* fake = nullptr;          // ASan fake-stack handle while switched out
    uint3 tid{};
    int lane = 0, warp = 0;
    bool done = false;
};
struct Warp {
    int live = 0, count = 0;
    unsigned gen = 0;
    uint32_t slot[32];
    uint32_t fresh = 0;            // lanes that wrote their slot in the running shuffle
};
struct Block {
    std::vector<Fiber> fib;
    std::vector<Warp> warps;
    int live = 0, bar_count = 0;
    unsigned bar_gen = 0;
    unsigned long progress = 0;
    int cur = 0;
    void* sched_sp = nullptr;
    const void* sched_bottom = nullptr;     // ASan: the scheduler's (host thread's) stack
    size_t sched_size = 0;
    void* sched_fake = nullptr;
    void* dsm = nullptr;
    const std::function<void()>* body = nullptr;
};
struct Sched {
    std::vector<uint8_t*> stacks;           // reused from launch to launch
    std::vector<uint8_t> dsm;
    Block blk;
    bool running = false;
    ~Sched() {
        for (uint8_t* s : stacks) munmap(s, STACK_BYTES);
    }
};
thread_local Sched tl_sched;
thread_local const char* tl_kernel = "";

void yield_to_scheduler(bool dying = false) {
    Block& b = tl_sched.blk;
#ifdef CUEMU_ASAN
    Fiber& f = b.fib[b.cur];
    __sanitizer_start_switch_fiber(dying ? nullptr : &f.fake, b.sched_bottom, b.sched_size);
    cuemu_switch(&f.sp, b.sched_sp);
    __sanitizer_finish_switch_fiber(f.fake, &b.sched_bottom, &b.sched_size);
#else
    (void)dying;
    cuemu_switch(&b.fib[b.cur].sp, b.sched_sp);
#endif
}

void release_block_barrier_if_complete(Block& b) {
    if (b.live > 0 && b.bar_count == b.live) {
        b.bar_count = 0;
        b.bar_gen++;
        b.progress++;
    }
}
void release_warp_barrier_if_complete(Block& b, Warp& w) {
    if (w.live > 0 && w.count == w.live) {
        w.count = 0;
        w.gen++;
        b.progress++;
    }
}

void warp_barrier() {
    Block& b = tl_sched.blk;
    Warp& w = b.warps[b.fib[b.cur].warp];
    const unsigned g = w.gen;
    w.count++;
    release_warp_barrier_if_complete(b, w);
    while (w.gen == g) yield_to_scheduler();
}

void fiber_main() {
    Block& b = tl_sched.blk;
#ifdef CUEMU_ASAN
    __sanitizer_finish_switch_fiber(nullptr, &b.sched_bottom, &b.sched_size);
#endif
    (*b.body)();
    Fiber& f = b.fib[b.cur];
    f.done = true;
    b.live--;
    b.progress++;
    Warp& w = b.warps[f.warp];
    w.live--;
    release_block_barrier_if_complete(b);       // the others may have been waiting for this thread only
    release_warp_barrier_if_complete(b, w);
    yield_to_scheduler(true);
    abort();                                    // a finished fiber is never resumed
}

void* fresh_context(uint8_t* stack) {
#ifdef CUEMU_ASAN
    __asan_unpoison_memory_region(stack, STACK_BYTES);      // the previous fiber on this stack never unwound its frames
#endif
    uintptr_t top = ((uintptr_t)stack + STACK_BYTES) & ~(uintptr_t)15;
    void** p = (void**)(top - 64);              // 16-byte aligned slot for the entry address: rsp % 16 == 8 at fiber_main's entry
    p[0] = (void*)&fiber_main;
    p[1] = nullptr;
    void** sp = p - 6;                          // rbp rbx r12 r13 r14 r15
    for (int i = 0; i < 6; i++) sp[i] = nullptr;
    return sp;
}
}  // namespace

// ---- operation census: how many Montgomery products / squarings each kernel executes (profiles/: the integer-multiply
// "roofline" of a proof, counted instead of estimated).  count_op is called from the field layer's host path (a line
// build_emu.py adds to the COPIED fp.cuh); only work done inside kernels is counted.
namespace {
struct OpCounts {
    unsigned long long launches = 0, threads = 0, mul[3] = {0, 0, 0}, sqr[3] = {0, 0, 0};      // [8, 12, 24 limbs]
};
thread_local OpCounts tl_ops;
std::mutex g_census_mu;
std::map<std::string, OpCounts> g_census;
void census_commit(const char* kernel, unsigned long long threads) {
    std::lock_guard<std::mutex> lk(g_census_mu);
    OpCounts& c = g_census[kernel];
    c.launches++;
    c.threads += threads;
    for (int i = 0; i < 3; i++) {
        c.mul[i] += tl_ops.mul[i];
        c.sqr[i] += tl_ops.sqr[i];
    }
    tl_ops = OpCounts();
}
}  // namespace
void count_op(int limbs, int squaring) {
    if (!tl_sched.running) return;
    const int k = limbs == 8 ? 0 : limbs == 12 ? 1 : 2;
    (squaring ? tl_ops.sqr : tl_ops.mul)[k]++;
}

int lane() { return tl_sched.blk.fib[tl_sched.blk.cur].lane; }
void* dyn_smem() { return tl_sched.blk.dsm; }

void sync_block() {
    Block& b = tl_sched.blk;
    const unsigned g = b.bar_gen;
    b.bar_count++;
    release_block_barrier_if_complete(b);
    while (b.bar_gen == g) yield_to_scheduler();
}
void sync_warp() { warp_barrier(); }
uint32_t shfl(uint32_t v, int src) {
    Block& b = tl_sched.blk;
    Fiber& f = b.fib[b.cur];
    Warp& w = b.warps[f.warp];
    w.slot[f.lane] = v;
    w.fresh |= 1u << f.lane;
    warp_barrier();
    // a source lane that takes no part in this shuffle (it returned earlier, or lies beyond the block's last thread) gives
    // UNDEFINED data on the hardware: poison instead of the stale slot, so a kernel that consumes such a value cannot pass here
    const bool taking_part = (w.fresh >> (src & 31)) & 1;
    const uint32_t r = taking_part ? w.slot[src & 31] : 0xBAADF00Du ^ (uint32_t)(b.progress * 2654435761u);
    if (!taking_part) {
        static const bool verbose = getenv("CUEMU_VERBOSE") != nullptr;
        static std::mutex mu;
        static std::map<std::string, int> seen;
        if (verbose) {
            std::lock_guard<std::mutex> lk(mu);
            if (seen[tl_kernel]++ == 0) fprintf(stderr, "cuemu: %s reads a shuffle value from a lane that takes no part (poisoned)\n", tl_kernel);
        }
    }
    warp_barrier();                             // nobody overwrites a slot before everybody has read
    w.fresh &= ~(1u << f.lane);
    return r;
}

void launch(dim3 grid, dim3 block, size_t dyn_bytes, const char* kernel, const std::function<void()>& body) {
    Sched& s = tl_sched;
    if (s.running) {
        fprintf(stderr, "cuemu: kernel launch from inside a kernel\n");
        abort();
    }
    const size_t nthreads = (size_t)block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > 1024 || block.z > 64 || (size_t)grid.x * grid.y * grid.z == 0 || grid.y > 65535 || grid.z > 65535 || dyn_bytes > (227u << 10)) {
        // what the hardware refuses (an empty grid or block, more than 1024 threads, too much shared memory): the launch does
        // not happen and cudaGetLastError() reports it, exactly as on the GPU
        tl_last_error = 9;       // cudaErrorInvalidConfiguration
        if (getenv("CUEMU_VERBOSE")) fprintf(stderr, "cuemu: invalid launch configuration for %s: grid (%u,%u,%u) block (%u,%u,%u) smem %zu\n", kernel, grid.x, grid.y, grid.z, block.x, block.y, block.z, dyn_bytes);
        return;
    }
    s.running = true;
    tl_kernel = kernel;
    while (s.stacks.size() < nthreads) {
        void* m = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) {
            perror("cuemu: mmap");
            abort();
        }
        s.stacks.push_back((uint8_t*)m);
    }
    if (s.dsm.size() < dyn_bytes + 64) s.dsm.resize(dyn_bytes + 64);
    Block& b = s.blk;
    b.body = &body;
    b.dsm = (void*)(((uintptr_t)s.dsm.data() + 63) & ~(uintptr_t)63);
    tl_gridDim = grid;
    tl_blockDim = block;
    const size_t nwarps = (nthreads + 31) / 32;
    static const int order_mode = [] {
        const char* e = getenv("CUEMU_ORDER");
        return !e ? 0 : !strcmp(e, "reverse") ? 1 : !strcmp(e, "shuffle") ? 2 : 0;
    }();
    // shuffle: t = (k * mul + add) mod n is a permutation of 0..n-1 whenever gcd(mul, n) = 1; blocks have 32..1024 threads
    // (powers of two times 1, 3 or 5), so any multiplier coprime to 2, 3 and 5 works
    const size_t order_mul = 7 * 11 * 13 + 0, order_add = 17;
    unsigned pass = 0;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                tl_blockIdx = uint3{bx, by, bz};
                b.fib.assign(nthreads, Fiber());
                b.warps.assign(nwarps, Warp());
                b.live = (int)nthreads;
                b.bar_count = 0;
                b.bar_gen = 0;
                b.progress = 0;
                for (size_t t = 0; t < nthreads; t++) {
                    Fiber& f = b.fib[t];
                    f.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / ((size_t)block.x * block.y))};
                    f.lane = (int)(t & 31);
                    f.warp = (int)(t >> 5);
                    f.sp = fresh_context(s.stacks[t]);
                    b.warps[f.warp].live++;
                }
                while (b.live > 0) {
                    const unsigned long before = b.progress;
                    for (size_t k = 0; k < nthreads; k++) {
                        // CUEMU_ORDER: the order in which the threads of a block get the CPU between rendezvous points.  A kernel
                        // without shared-memory / global-memory hazards gives the same bytes for any order; running the tests
                        // with "reverse" and "shuffle" is the emulator's (weak) counterpart of a race check.
                        const size_t t = order_mode == 0 ? k : order_mode == 1 ? nthreads - 1 - k : (size_t)((k * order_mul + order_add + pass * 7919u) % nthreads);
                        if (b.fib[t].done) continue;
                        b.cur = (int)t;
                        tl_threadIdx = b.fib[t].tid;
#ifdef CUEMU_ASAN
                        __sanitizer_start_switch_fiber(&b.sched_fake, s.stacks[t], STACK_BYTES);
                        cuemu_switch(&b.sched_sp, b.fib[t].sp);
                        __sanitizer_finish_switch_fiber(b.sched_fake, nullptr, nullptr);
#else
                        cuemu_switch(&b.sched_sp, b.fib[t].sp);
#endif
                    }
                    pass++;
                    if (b.live > 0 && b.progress == before) {
                        fprintf(stderr, "cuemu: deadlock in block (%u,%u,%u): %d threads wait at a rendezvous the others never reach\n", bx, by, bz, b.live);
                        abort();
                    }
                }
            }
    s.running = false;
    census_commit(kernel, (unsigned long long)grid.x * grid.y * grid.z * nthreads);
}

}  // namespace cuemu

// ------------------------------------------------------------------------------------------------- runtime API
namespace {
thread_local int tl_device = 0;
int device_count() {
    const char* e = getenv("CUEMU_DEVICES");
    int n = e ? atoi(e) : 1;
    return n < 1 ? 1 : n > 16 ? 16 : n;
}
std::atomic<long> g_live_allocs{0};
}  // namespace

extern "C" {
cudaError_t cudaGetDeviceCount(int* n) {
    *n = device_count();
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int d) {
    if (d < 0 || d >= device_count()) return cudaErrorInvalidDevice;
    tl_device = d;
    return cudaSuccess;
}
cudaError_t cudaGetDevice(int* d) {
    *d = tl_device;
    return cudaSuccess;
}
cudaError_t cudaMalloc(void** p, size_t n) {
    void* m = nullptr;
    if (n == 0) {                 // the runtime hands back a null pointer for an empty allocation
        *p = nullptr;
        return cudaSuccess;
    }
    if (posix_memalign(&m, 256, n) != 0) {
        *p = nullptr;
        return cudaErrorMemoryAllocation;
    }
    // device memory is not zeroed: make accidental reliance on its contents visible.  CUEMU_FILL=<byte> fills every allocation
    // completely with that byte (run the tests with 0, 255, ...: results must not depend on it -- a weak initcheck)
    static const int fill = [] {
        const char* e = getenv("CUEMU_FILL");
        return e ? atoi(e) & 255 : -1;
    }();
    if (fill >= 0) memset(m, fill, n);
    else memset(m, 0xA5, n < 4096 ? n : 4096);
    g_live_allocs++;
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    if (p) g_live_allocs--;
    free(p);
    return cudaSuccess;
}
__attribute__((visibility("default"))) long cuemu_live_allocations(void) { return g_live_allocs.load(); }
__attribute__((visibility("default"))) void cuemu_census_reset(void) {
    std::lock_guard<std::mutex> lk(cuemu::g_census_mu);
    cuemu::g_census.clear();
}
// JSON object kernel -> {launches, threads, mul8, sqr8, mul12, sqr12, mul24, sqr24}; "mul" counts every product, the
// squarings among them are listed again under "sqr".  Returns the length needed (call again with a larger buffer).
__attribute__((visibility("default"))) size_t cuemu_census_json(char* buf, size_t cap) {
    std::lock_guard<std::mutex> lk(cuemu::g_census_mu);
    std::string o = "{";
    bool first = true;
    for (auto& kv : cuemu::g_census) {
        char line[512];
        const auto& c = kv.second;
        snprintf(line, sizeof line, "%s\"%s\": {\"launches\": %llu, \"threads\": %llu, \"mul8\": %llu, \"sqr8\": %llu, \"mul12\": %llu, \"sqr12\": %llu, \"mul24\": %llu, \"sqr24\": %llu}",
                 first ? "" : ", ", kv.first.c_str(), c.launches, c.threads, c.mul[0], c.sqr[0], c.mul[1], c.sqr[1], c.mul[2], c.sqr[2]);
        o += line;
        first = false;
    }
    o += "}";
    if (buf && cap > o.size()) memcpy(buf, o.c_str(), o.size() + 1);
    return o.size() + 1;
}
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
    void* m = nullptr;
    if (posix_memalign(&m, 256, n ? n : 256) != 0) return cudaErrorMemoryAllocation;
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) {
    free(p);
    return cudaSuccess;
}
cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t) {
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
    if (n) memset(d, v, n);
    return cudaSuccess;
}
cudaError_t cudaMemset(void* d, int v, size_t n) {
    if (n) memset(d, v, n);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
    *s = (cudaStream_t)malloc(8);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) {
    *s = (cudaStream_t)malloc(8);
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) {
    free(s);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) {
    if (least) *least = 0;
    if (greatest) *greatest = -1;
    return cudaSuccess;
}
cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) {
    *can = 1;
    return cudaSuccess;
}
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) {
    *e = (cudaEvent_t)malloc(8);
    return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
    *e = (cudaEvent_t)malloc(8);
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e) {
    free(e);
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) {
    *ms = 0.f;
    return cudaSuccess;
}
cudaError_t cudaGetLastError(void) {
    cudaError_t e = cuemu::tl_last_error;
    cuemu::tl_last_error = 0;
    return e;
}
cudaError_t cudaPeekAtLastError(void) { return cuemu::tl_last_error; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == 9 ? "invalid configuration argument" : "emulated CUDA error"; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    a->type = cudaMemoryTypeUnregistered;          // every host buffer is "pageable": the library's staging ring gets exercised
    a->device = 0;
    a->devicePointer = nullptr;
    a->hostPointer = (void*)p;
    return cudaSuccess;
}
}
