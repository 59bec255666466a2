// tests/host/cuemu/selftest.cpp -- TEST SCAFFOLDING: the emulator checked on kernels whose answer is known, so that "the suite
// passes on the emulator" means something.  Prints one "name value" line per check (tests/test_emulated_library.py reads them).
#include "cuda_runtime.h"

#include <cstdio>
#include <tuple>
#include <vector>

// every lane asks for the value of lane (l + 8) % 32; lanes 24..31 returned before the shuffle, the block has 40 threads
// (second warp: 8 lanes).  On the GPU a value read from a lane that takes no part is undefined: the emulator must not hand
// back that lane's stale slot, it poisons the value.
static void k_shuffle(uint32_t* out) {
    const unsigned t = threadIdx.x, l = t & 31;
    uint32_t first = __shfl_sync(0xFFFFFFFFu, 1000 + t, (l + 1) & 31);      // everybody takes part: fills every slot
    if (t < 32 && l >= 24) {
        out[t] = first;
        return;
    }
    out[t] = __shfl_sync(0xFFFFFFFFu, 2000 + t, (l + 8) & 31);
}

// a block-wide exchange through shared memory with a barrier on both sides: thread t reads what thread n-1-t wrote
static void k_exchange(uint32_t* out) {
    __shared__ uint32_t buf[96];
    const unsigned t = threadIdx.x, n = blockDim.x;
    for (int round = 0; round < 3; round++) {
        buf[t] = 100 * round + t + blockIdx.x;
        __syncthreads();
        const uint32_t v = buf[n - 1 - t];
        __syncthreads();
        out[(blockIdx.x * 3 + round) * n + t] = v;
    }
}

int main() {
    std::vector<uint32_t> o(40, 7);
    uint32_t* p = o.data();
    cuemu::launch(dim3(1), dim3(40), 0, "k_shuffle", [&]() { k_shuffle(p); });
    int good = 0, poisoned = 0, stale = 0;
    for (unsigned t = 0; t < 40; t++) {
        const unsigned l = t & 31, src = (l + 8) & 31;
        if (t < 32 && l >= 24) {
            good += o[t] == 1000 + ((l + 1) & 31);
            continue;
        }
        const bool src_takes_part = t < 32 ? src < 24 : src < 8;
        if (src_takes_part)
            good += o[t] == 2000 + (t & ~31u) + src;
        else {
            stale += o[t] == 1000 + (t & ~31u) + src || o[t] == 2000 + (t & ~31u) + src;
            poisoned += (o[t] >> 16) != 0;
        }
    }
    printf("shuffle_good %d\nshuffle_poisoned %d\nshuffle_stale %d\n", good, poisoned, stale);

    std::vector<uint32_t> e(2 * 3 * 96, 0);
    uint32_t* q = e.data();
    cuemu::launch(dim3(2), dim3(96), 0, "k_exchange", [&]() { k_exchange(q); });
    int ok = 0;
    for (unsigned b = 0; b < 2; b++)
        for (unsigned r = 0; r < 3; r++)
            for (unsigned t = 0; t < 96; t++) ok += e[(b * 3 + r) * 96 + t] == 100 * r + (95 - t) + b;
    printf("exchange_good %d\n", ok);

    // ---- streams.  Two mis-ordered programs and their repaired versions; what each prints depends on CUEMU_ASYNC:
    //   missing wait   the consumer (stream B) copies a buffer the producer (stream A) fills, with no event between them:
    //                  right in the synchronous mode and under `eager`, STALE under `lazy` (B runs before A)
    //   early reuse    A refills the buffer without waiting for B's read of the first contents:
    //                  right in the synchronous mode and under `lazy`, OVERWRITTEN under `eager` (A runs to its end before B)
    // the repaired versions (event record + wait) give the right bytes in every mode
    for (int repaired = 0; repaired < 2; repaired++) {
        uint8_t *a, *b;
        cudaMalloc((void**)&a, 64);
        cudaMalloc((void**)&b, 64);
        cudaStream_t A, B;
        cudaStreamCreateWithFlags(&A, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&B, cudaStreamNonBlocking);
        cudaEvent_t filled, read;
        cudaEventCreate(&filled);
        cudaEventCreate(&read);
        cudaMemsetAsync(a, 0, 64, A);
        cudaMemsetAsync(b, 0, 64, B);
        cudaStreamSynchronize(A);
        cudaStreamSynchronize(B);
        // missing wait
        cudaMemsetAsync(a, 1, 64, A);
        if (repaired) {
            cudaEventRecord(filled, A);
            cudaStreamWaitEvent(B, filled, 0);
        }
        cudaMemcpyAsync(b, a, 64, cudaMemcpyDeviceToDevice, B);
        cudaStreamSynchronize(B);
        const int got1 = b[0];
        cudaStreamSynchronize(A);
        // early reuse
        cudaMemsetAsync(a, 5, 64, A);
        cudaEventRecord(filled, A);
        cudaStreamWaitEvent(B, filled, 0);
        cudaMemcpyAsync(b, a, 64, cudaMemcpyDeviceToDevice, B);
        if (repaired) {
            cudaEventRecord(read, B);
            cudaStreamWaitEvent(A, read, 0);
        }
        cudaMemsetAsync(a, 9, 64, A);
        cudaStreamSynchronize(B);
        const int got2 = b[0];
        cudaDeviceSynchronize();
        printf("%s_consumer_sees %d\n%s_reader_sees %d\n", repaired ? "ordered" : "unordered", got1, repaired ? "ordered" : "unordered", got2);
        cudaStreamDestroy(A);
        cudaStreamDestroy(B);
        cudaEventDestroy(filled);
        cudaEventDestroy(read);
        cudaFree(a);
        cudaFree(b);
    }
    // a launch's arguments are evaluated at the launch, not when the kernel runs
    {
        uint32_t* d;
        cudaMalloc((void**)&d, 4 * 40);
        cudaStream_t S;
        cudaStreamCreateWithFlags(&S, cudaStreamNonBlocking);
        uint32_t* target = d;
        { auto args = std::make_tuple(target); cuemu::launch_on(S, dim3(1), dim3(40), 0, "k_shuffle", [=]() { std::apply([=](auto&... x) { k_shuffle(x...); }, args); }); }
        target = nullptr;                      // must not matter any more
        cudaStreamSynchronize(S);
        printf("launch_arguments_kept %d\n", d[0] == 2008 ? 1 : 0);
        cudaStreamDestroy(S);
        cudaFree(d);
    }

    // what the hardware refuses: an empty grid, 1025 threads
    cuemu::launch(dim3(0), dim3(32), 0, "none", [&]() {});
    const int e1 = (int)cudaGetLastError();
    cuemu::launch(dim3(1), dim3(1025), 0, "none", [&]() {});
    const int e2 = (int)cudaGetLastError(), e3 = (int)cudaGetLastError();
    printf("empty_grid_error %d\ntoo_many_threads_error %d\nerror_cleared %d\n", e1, e2, e3);
    return 0;
}
