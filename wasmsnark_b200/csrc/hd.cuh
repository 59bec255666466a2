// hd.cuh -- host/device qualifiers shared by the arithmetic headers.
// The field / curve headers compile both with nvcc (device path: PTX carry chains)
// and with plain g++ (host path: portable 64-bit C) so that the group-law logic can be
// unit-tested on a CPU-only box; the host path is test scaffolding, never a fallback --
// the library's entry points only ever launch the device kernels.
#pragma once
#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#define HD_NOINLINE __host__ __device__ __noinline__
#else
#define HD inline
#define DEV inline
#define HD_NOINLINE __attribute__((noinline))
#endif
