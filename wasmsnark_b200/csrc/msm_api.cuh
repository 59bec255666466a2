// msm_api.cuh -- non-template MSM entry points shared between translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
struct ws_ctx;
namespace ws {
// Enqueue sum s_i P_i on `st`.  d_out_jac (96 / 192 bytes) and d_out_xyzz (128 / 256 bytes) may each
// be null.  arena_slot selects the ctx workspace arena (concurrent MSMs need distinct slots).
// Defined in msm_g1.cu / msm_g2.cu.
int msm_g1_run(ws_ctx* c, const void* d_scalars, const void* d_points, size_t n, void* d_out_jac, void* d_out_xyzz, cudaStream_t st, int arena_slot = 1);
int msm_g2_run(ws_ctx* c, const void* d_scalars, const void* d_points, size_t n, void* d_out_jac, void* d_out_xyzz, cudaStream_t st, int arena_slot = 1);
}  // namespace ws
