// Shared declarations for the NVP hot-path kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/nvp_b200.h"

namespace nvp {

constexpr int kHidden = 128;

// Thread-local error string + launch counter behind nvp_last_error()/nvp_last_launch_count().
void set_error(const std::string& msg);
void count_launch(int n = 1);
void reset_launch_count();
// Event registered with nvp_record_grid_grads_event for the current backward call (cleared by the read), or nullptr.
cudaEvent_t take_grid_event();

#define NVP_CHECK(cond, msg)                                   \
  do {                                                         \
    if (!(cond)) {                                             \
      ::nvp::set_error(std::string(msg));                      \
      return 1;                                                \
    }                                                          \
  } while (0)

#define NVP_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::nvp::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
      return 2;                                                                             \
    }                                                                                       \
  } while (0)

#define NVP_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    ::nvp::count_launch();                                                                  \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      ::nvp::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e));     \
      return 3;                                                                             \
    }                                                                                       \
  } while (0)

// Optional per-kernel CUDA-event timing (nvp_profile_enable / nvp_profile_read in the C ABI).
enum KernelId { K_PACK = 0, K_GATHER, K_MLP_FWD, K_MLP_BWD, K_MLP_WGRAD, K_SCATTER, K_SIMT, K_MISC, K_BIN, K_MLP_FUSED, K_COUNT };
void prof_start(int id, cudaStream_t st);
void prof_stop(cudaStream_t st);
struct ScopedKernelTimer {
  cudaStream_t st;
  ScopedKernelTimer(int id, cudaStream_t s) : st(s) { prof_start(id, s); }
  ~ScopedKernelTimer() { prof_stop(st); }
};

// DenseGrid level layout (device-visible copy passed by value to the grid kernels).
struct LevelTab {
  float scale[NVP_MAX_LEVELS];
  int32_t res[NVP_MAX_LEVELS];
  int32_t offset[NVP_MAX_LEVELS + 1];  // in cells
  int32_t n_levels;
};

int build_level_table(const nvp_desc* d, LevelTab* tab, int64_t* offsets64);
int validate_desc(const nvp_desc* d);

inline int latent_dim(const nvp_desc* d) { return 3 * d->n_levels * d->n_features + 9 * d->sparse_features; }
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---- grid.cu ------------------------------------------------------------------------------
// z (fp32 row-major, pitch ldz floats) and/or z16t (fp16 MMA tile format, kz 64-column panels per
// 128-sample tile; see tc_common.cuh).
int launch_grid_gather(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords,
                       int64_t n, float* z, int ldz, uint8_t* z16t, int kz, cudaStream_t st,
                       bool temporal_interp = false);
// grads += scatter of dz scaled by `scale` (times *scale_ptr when given).  dz is either fp32 row-major
// (pitch lddz) or, when dz16t != NULL, fp16 in the MMA tile format with kz panels per 128-sample tile.
int launch_grid_scatter(const nvp_desc* d, const LevelTab& tab, const float* coords, int64_t n,
                        const float* dz, int lddz, const uint8_t* dz16t, int kz, float scale, const float* scale_ptr,
                        const nvp_grads* g, cudaStream_t st);

// Tile-binned variant for the tensor-core path (grid_binned.cuh): samples are bucketed per keyframe plane by the tile
// of the unit square they fall into (launch_grid_bin, once per call) and the gather / scatter-add work on private
// shared-memory windows.  grid_bin_workspace_bytes == 0 means "not available for this configuration".
size_t grid_bin_workspace_bytes(const nvp_desc* d, const LevelTab& tab, int64_t n);
// Host-side description of the plan (nvp_grid_bin_plan); tb = 0: not binned.  extent/base may be nullptr.
void grid_bin_plan_info(const nvp_desc* d, const LevelTab& tab, int64_t n, int* tb, int* chunk, int32_t* extent, int32_t* base,
                        size_t* workspace);
int launch_grid_bin(const nvp_desc* d, const LevelTab& tab, const float* coords, int64_t n, int kz, void* binws,
                    cudaStream_t st);
int launch_grid_gather_binned(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords,
                              int64_t n, uint8_t* z16t, int kz, void* binws, cudaStream_t st,
                              bool temporal_interp = false);
int launch_grid_scatter_binned(const nvp_desc* d, const LevelTab& tab, const float* coords, int64_t n,
                               const uint8_t* dz16t, int kz, float scale, const float* scale_ptr, const nvp_grads* g,
                               void* binws, cudaStream_t st);

// ---- mlp_simt.cu --------------------------------------------------------------------------
size_t simt_workspace_bytes(const nvp_desc* d, int64_t n, int what);
int simt_forward(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords,
                 const float* tsteps, int64_t n, float* out_rgb, void* ws, size_t ws_bytes, cudaStream_t st,
                 bool temporal_interp = false);
// dout == nullptr -> loss mode (gt_u8, n_global, loss_sum used); else explicit upstream gradient.
int simt_fwd_bwd(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords,
                 const float* tsteps, const uint8_t* gt_u8, const float* dout, int64_t n, int64_t n_global,
                 const nvp_grads* g, float* loss_sum, float* out_rgb, void* ws, size_t ws_bytes, cudaStream_t st);

// ---- mlp_tc.cu ----------------------------------------------------------------------------
size_t tc_workspace_bytes(const nvp_desc* d, int64_t n, int what);
int tc_timeline_read(unsigned long long* out, int n);   // NVP_TIMELINE builds only
int tc_forward(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords,
               const float* tsteps, int64_t n, float* out_rgb, void* ws, size_t ws_bytes, cudaStream_t st,
               bool temporal_interp = false);
int tc_fwd_bwd(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords,
               const float* tsteps, const uint8_t* gt_u8, const float* dout, int64_t n, int64_t n_global,
               const nvp_grads* g, float* loss_sum, float* out_rgb, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace nvp
