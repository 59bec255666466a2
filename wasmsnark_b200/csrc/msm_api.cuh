// msm_api.cuh -- non-template MSM entry points shared between translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
struct ws_ctx;
namespace ws {
// Enqueue sum s_i P_i on `st`.  d_out_jac (96 / 192 bytes) and d_out_xyzz (128 / 256 bytes) may each
// be null.  arena_slot selects the ctx workspace arena (concurrent MSMs need distinct slots).
// Defined in msm_g1.cu / msm_g2.cu.
int msm_g1_run(ws_ctx* c, const void* d_scalars, const void* d_points, size_t n, void* d_out_jac, void* d_out_xyzz, cudaStream_t st, int arena_slot = 1, int table_c = 0);
int msm_g2_run(ws_ctx* c, const void* d_scalars, const void* d_points, size_t n, void* d_out_jac, void* d_out_xyzz, cudaStream_t st, int arena_slot = 1, int table_c = 0);
// Window tables for resident keys: table_c = window bits; the table holds msm_table_windows_n(table_c) * n affine
// points.  With table_c != 0 in msm_g?_run, d_points must be such a table.
int msm_table_windows_n(int table_c);
int msm_table_choose_c(size_t n);
int msm_g1_build_table(ws_ctx* c, const void* d_points, size_t n, int table_c, void* d_table, cudaStream_t st);
int msm_g2_build_table(ws_ctx* c, const void* d_points, size_t n, int table_c, void* d_table, cudaStream_t st);
}  // namespace ws
