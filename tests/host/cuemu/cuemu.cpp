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

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
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

bool bad_launch_configuration(dim3 grid, dim3 block, size_t dyn_bytes) {
    const size_t nthreads = (size_t)block.x * block.y * block.z;
    return nthreads == 0 || nthreads > 1024 || block.z > 64 || (size_t)grid.x * grid.y * grid.z == 0 || grid.y > 65535 || grid.z > 65535 || dyn_bytes > (227u << 10);
}

void launch(dim3 grid, dim3 block, size_t dyn_bytes, const char* kernel, const std::function<void()>& body) {
    Sched& s = tl_sched;
    if (s.running) {
        fprintf(stderr, "cuemu: kernel launch from inside a kernel\n");
        abort();
    }
    const size_t nthreads = (size_t)block.x * block.y * block.z;
    if (bad_launch_configuration(grid, block, dyn_bytes)) {
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
// Two modes.  Default: synchronous -- every launch and copy happens inside the call that asks for it.
// CUEMU_ASYNC=lazy|eager: DEFERRED -- launches and asynchronous copies are queued on their stream and run only when
// something forces them: a synchronize, a blocking copy, an event another forced stream waits for.  `lazy` then runs the
// bare minimum (the forced stream, and of other streams only what its event waits name); `eager` runs every OTHER stream to
// the end first.  Both are legal schedules of the same program; a program whose streams are ordered by stream order and
// events alone gives the same bytes under all three.  One that relies on timing -- a consumer without a wait on its
// producer's event, a buffer reused before its last reader was ordered behind the writer -- gives different bytes under
// lazy (consumers run before unordered producers) or eager (unordered later writers run before earlier readers).  That is
// the class of defect a GPU run hides by luck and the synchronous emulator hides by construction.
namespace {
thread_local int tl_device = 0;
int device_count() {
    const char* e = getenv("CUEMU_DEVICES");
    int n = e ? atoi(e) : 1;
    return n < 1 ? 1 : n > 16 ? 16 : n;
}
std::atomic<long> g_live_allocs{0};

int async_mode() {
    static const int m = [] {
        const char* e = getenv("CUEMU_ASYNC");
        return !e ? 0 : !strcmp(e, "lazy") ? 1 : !strcmp(e, "eager") ? 2 : !strcmp(e, "random") ? 3 : 0;
    }();
    return m;
}
// CUEMU_ASYNC=random (CUEMU_SEED=<n>): at every forcing point ONE operation at a time, from a randomly chosen stream whose head
// can run (a wait whose event is not there yet blocks its stream), until what was forced is done -- interleavings between
// the two extremes, at the granularity of single launches and copies
unsigned long long g_rng = [] {
    const char* e = getenv("CUEMU_SEED");
    return 0x9E3779B97F4A7C15ull * (unsigned long long)(e ? atoll(e) + 1 : 1) | 1;
}();
unsigned next_random() {
    g_rng ^= g_rng << 13;
    g_rng ^= g_rng >> 7;
    g_rng ^= g_rng << 17;
    return (unsigned)(g_rng >> 24);
}
std::recursive_mutex g_rt_mu;
struct Ev {
    unsigned long long enqueued = 0, done = 0;      // record calls issued / executed
};
struct Op {
    int kind;                                       // 0 work, 1 record, 2 wait
    std::function<void()> fn;
    std::shared_ptr<Ev> ev;
    unsigned long long gen;
};
struct Stream {
    std::deque<Op> q;
    bool nonblocking = true, running = false;
    unsigned long long id = 0;
};
unsigned long long g_next_stream_id = 1;
std::map<cudaStream_t, std::shared_ptr<Stream>> g_streams;          // live handles
std::vector<std::shared_ptr<Stream>> g_zombies;                     // destroyed with work pending: it still completes
std::map<cudaEvent_t, std::shared_ptr<Ev>> g_events;
std::map<uintptr_t, size_t> g_device_ranges, g_pinned_ranges;
unsigned long long g_deferred_ops = 0, g_forced_drains = 0;

bool in_ranges(const std::map<uintptr_t, size_t>& m, const void* p) {
    auto it = m.upper_bound((uintptr_t)p);
    if (it == m.begin()) return false;
    --it;
    return (uintptr_t)p < it->first + it->second;
}
std::shared_ptr<Stream> stream_of(cudaStream_t h) {                  // h != nullptr
    auto it = g_streams.find(h);
    if (it != g_streams.end()) return it->second;
    auto s = std::make_shared<Stream>();                             // a handle the runtime did not hand out (tests): its own queue
    s->id = g_next_stream_id++;
    return g_streams[h] = s;
}
std::vector<std::shared_ptr<Stream>> all_streams() {
    std::vector<std::shared_ptr<Stream>> v;
    for (auto& kv : g_streams) v.push_back(kv.second);
    for (auto& z : g_zombies) v.push_back(z);
    std::sort(v.begin(), v.end(), [](const std::shared_ptr<Stream>& a, const std::shared_ptr<Stream>& b) { return a->id < b->id; });
    return v;
}
void run_stream(Stream& s, const Ev* until_ev, unsigned long long until_gen);
// random mode: single steps of randomly chosen runnable streams until `done()` holds
template <class Done>
void random_steps(Done done) {
    while (!done()) {
        std::vector<std::shared_ptr<Stream>> can;
        for (auto& sp : all_streams())
            if (!sp->q.empty() && (sp->q.front().kind != 2 || sp->q.front().ev->done >= sp->q.front().gen)) can.push_back(sp);
        if (can.empty()) {
            fprintf(stderr, "cuemu: nothing can run but something is waited for (a wait on an event whose record is stuck behind it)\n");
            abort();
        }
        Stream& s = *can[next_random() % can.size()];
        Op op = std::move(s.q.front());
        s.q.pop_front();
        if (op.kind == 0) op.fn();
        else if (op.kind == 1) op.ev->done = std::max(op.ev->done, op.gen);
    }
}

// make `ev` reach generation `gen`: run the stream that holds that record, up to it
void advance_event(const std::shared_ptr<Ev>& ev, unsigned long long gen) {
    if (async_mode() == 3) {
        random_steps([&] { return ev->done >= gen; });
        return;
    }
    while (ev->done < gen) {
        bool found = false;
        for (auto& sp : all_streams()) {
            for (auto& op : sp->q)
                if (op.kind == 1 && op.ev == ev && op.gen >= gen) {
                    found = true;
                    break;
                }
            if (found) {
                if (sp->running) {
                    fprintf(stderr, "cuemu: circular wait between streams (an event is awaited by work queued in front of its own record)\n");
                    abort();
                }
                run_stream(*sp, ev.get(), gen);
                break;
            }
        }
        if (!found) {
            fprintf(stderr, "cuemu: wait for an event record that is in no queue\n");
            abort();
        }
    }
}
// execute the head of `s` until its queue is empty, or (until_ev) until that event has reached until_gen
void run_stream(Stream& s, const Ev* until_ev, unsigned long long until_gen) {
    s.running = true;
    while (!s.q.empty() && !(until_ev && until_ev->done >= until_gen)) {
        Op op = std::move(s.q.front());
        s.q.pop_front();
        if (op.kind == 0) op.fn();
        else if (op.kind == 1) op.ev->done = std::max(op.ev->done, op.gen);
        else advance_event(op.ev, op.gen);
    }
    s.running = false;
}
void reap_zombies() {
    g_zombies.erase(std::remove_if(g_zombies.begin(), g_zombies.end(), [](const std::shared_ptr<Stream>& z) { return z->q.empty(); }), g_zombies.end());
}
void drain_everything(const Stream* last = nullptr) {
    if (async_mode() == 3) {
        random_steps([&] {
            for (auto& sp : all_streams())
                if (sp.get() != last && !sp->q.empty()) return false;
            return true;
        });
        reap_zombies();
        return;
    }
    for (auto& sp : all_streams())
        if (sp.get() != last && !sp->running) run_stream(*sp, nullptr, 0);
    reap_zombies();
}
// a forcing point on one stream
void force_stream(const std::shared_ptr<Stream>& s) {
    g_forced_drains++;
    if (async_mode() == 3) {
        random_steps([&] { return s->q.empty(); });
        reap_zombies();
        return;
    }
    if (async_mode() == 2) drain_everything(s.get());              // eager: everybody else first, to the end
    if (!s->running) run_stream(*s, nullptr, 0);
    reap_zombies();
}
// the legacy default stream orders itself against the BLOCKING streams only
void force_blocking_streams() {
    g_forced_drains++;
    if (async_mode() == 2) {
        drain_everything();
        return;
    }
    if (async_mode() == 3) {
        random_steps([&] {
            for (auto& sp : all_streams())
                if (!sp->nonblocking && !sp->q.empty()) return false;
            return true;
        });
        return;
    }
    for (auto& sp : all_streams())
        if (!sp->nonblocking && !sp->running) run_stream(*sp, nullptr, 0);
}
void enqueue(cudaStream_t h, std::function<void()> fn) {
    if (!h) {                                                       // legacy default stream: after the blocking streams, now
        force_blocking_streams();
        fn();
        return;
    }
    g_deferred_ops++;
    stream_of(h)->q.push_back(Op{0, std::move(fn), nullptr, 0});
}
cudaStream_t new_stream(unsigned flags) {
    cudaStream_t h = (cudaStream_t)malloc(8);
    if (async_mode()) {
        std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
        auto s = std::make_shared<Stream>();
        s->nonblocking = (flags & cudaStreamNonBlocking) != 0;
        s->id = g_next_stream_id++;
        g_streams[h] = s;
    }
    return h;
}
cudaError_t copy_async(void* d, const void* s, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    if (!async_mode()) {
        memmove(d, s, n);
        return cudaSuccess;
    }
    std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
    const bool src_dev = in_ranges(g_device_ranges, s), dst_dev = in_ranges(g_device_ranges, d);
    const bool src_pin = in_ranges(g_pinned_ranges, s), dst_pin = in_ranges(g_pinned_ranges, d);
    if (!src_dev && !dst_dev) {                                     // host to host: done by the call
        memmove(d, s, n);
        return cudaSuccess;
    }
    if (!src_dev && !src_pin) {                                     // from pageable memory: staged before the call returns
        auto snap = std::make_shared<std::vector<uint8_t>>((const uint8_t*)s, (const uint8_t*)s + n);
        enqueue(st, [d, snap] { memcpy(d, snap->data(), snap->size()); });
        return cudaSuccess;
    }
    if (!dst_dev && !dst_pin) {                                     // into pageable memory: the call returns when the copy is done
        if (st) force_stream(stream_of(st));
        else force_blocking_streams();
        memmove(d, s, n);
        return cudaSuccess;
    }
    enqueue(st, [d, s, n] { memmove(d, s, n); });                   // device / pinned on both sides: truly asynchronous
    return cudaSuccess;
}
}  // namespace

namespace cuemu {
void launch_on(cudaStream_t st, dim3 grid, dim3 block, size_t dyn_bytes, const std::string& kernel, std::function<void()> body) {
    if (!async_mode()) {
        launch(grid, block, dyn_bytes, kernel.c_str(), body);
        return;
    }
    if (bad_launch_configuration(grid, block, dyn_bytes)) {         // refused at the launch, as on the GPU
        tl_last_error = 9;
        return;
    }
    std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
    enqueue(st, [grid, block, dyn_bytes, kernel, body] { launch(grid, block, dyn_bytes, kernel.c_str(), body); });
}
}  // namespace cuemu

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
    if (async_mode()) {
        std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
        g_device_ranges[(uintptr_t)m] = n;
    }
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    if (p && async_mode()) {                      // cudaFree synchronises the device before it releases anything
        std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
        drain_everything();
        g_device_ranges.erase((uintptr_t)p);
    }
    if (p) g_live_allocs--;
    free(p);
    return cudaSuccess;
}
__attribute__((visibility("default"))) int cuemu_async_mode(void) { return async_mode(); }
// everything queued anywhere runs now (the tests' torch stand-ins: torch.cuda.synchronize(), tensor.cpu(), a collective)
__attribute__((visibility("default"))) void cuemu_sync_all(void) {
    if (!async_mode()) return;
    std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
    drain_everything();
}
__attribute__((visibility("default"))) unsigned long long cuemu_deferred_ops(void) { return g_deferred_ops; }
// launches / copies / event records and waits sitting in the queues right now
__attribute__((visibility("default"))) unsigned long long cuemu_pending_ops(void) {
    std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
    unsigned long long n = 0;
    for (auto& sp : all_streams()) n += sp->q.size();
    return n;
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
    if (async_mode()) {
        std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
        g_pinned_ranges[(uintptr_t)m] = n ? n : 256;
    }
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) {
    if (p && async_mode()) {
        std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
        drain_everything();
        g_pinned_ranges.erase((uintptr_t)p);
    }
    free(p);
    return cudaSuccess;
}
cudaError_t cudaHostRegister(void* p, size_t n, unsigned) {
    if (async_mode()) {
        std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
        g_pinned_ranges[(uintptr_t)p] = n;
    }
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    if (async_mode()) {                           // a blocking copy: the legacy default stream's ordering, then the copy itself
        std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
        force_blocking_streams();
    }
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t st) { return copy_async(d, s, n, st); }
cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t st) { return copy_async(d, s, n, st); }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    if (!async_mode()) {
        memset(d, v, n);
        return cudaSuccess;
    }
    std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
    enqueue(st, [d, v, n] { memset(d, v, n); });
    return cudaSuccess;
}
cudaError_t cudaMemset(void* d, int v, size_t n) {
    if (async_mode()) {
        std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
        force_blocking_streams();
    }
    if (n) memset(d, v, n);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags) {
    *s = new_stream(flags);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned flags, int) {
    *s = new_stream(flags);
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) {
    if (async_mode()) {                           // returns at once; the work queued on the stream still completes
        std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
        auto it = g_streams.find(s);
        if (it != g_streams.end()) {
            if (!it->second->q.empty()) g_zombies.push_back(it->second);
            g_streams.erase(it);
        }
    }
    free(s);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) {
    if (!async_mode()) return cudaSuccess;
    std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
    if (s) force_stream(stream_of(s));
    else force_blocking_streams();
    return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
    if (!async_mode()) return cudaSuccess;
    std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
    auto it = g_events.find(e);
    if (it == g_events.end() || it->second->enqueued == 0) return cudaSuccess;         // never recorded: nothing to wait for
    if (!s) {
        force_blocking_streams();
        advance_event(it->second, it->second->enqueued);
        return cudaSuccess;
    }
    stream_of(s)->q.push_back(Op{2, nullptr, it->second, it->second->enqueued});       // the most recent record at THIS moment
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize(void) {
    if (!async_mode()) return cudaSuccess;
    std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
    g_forced_drains++;
    drain_everything();
    return cudaSuccess;
}
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
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
    *e = (cudaEvent_t)malloc(8);
    if (async_mode()) {
        std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
        g_events[*e] = std::make_shared<Ev>();
    }
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { return cudaEventCreateWithFlags(e, 0); }
cudaError_t cudaEventDestroy(cudaEvent_t e) {
    if (async_mode()) {                           // queued records / waits keep the state alive
        std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
        g_events.erase(e);
    }
    free(e);
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
    if (!async_mode()) return cudaSuccess;
    std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
    auto it = g_events.find(e);
    if (it == g_events.end()) return cudaErrorInvalidValue;
    const unsigned long long gen = ++it->second->enqueued;
    if (!s) {
        force_blocking_streams();
        it->second->done = gen;
        return cudaSuccess;
    }
    stream_of(s)->q.push_back(Op{1, nullptr, it->second, gen});
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) {
    if (!async_mode()) return cudaSuccess;
    std::lock_guard<std::recursive_mutex> lk(g_rt_mu);
    auto it = g_events.find(e);
    if (it == g_events.end()) return cudaErrorInvalidValue;
    g_forced_drains++;
    if (async_mode() == 2) drain_everything();
    advance_event(it->second, it->second->enqueued);
    return cudaSuccess;
}
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
