// tcgen05 tensor-core path (placeholder until the kernels land).
#include "common.cuh"
namespace nvp {
size_t tc_workspace_bytes(const nvp_desc*, int64_t, int) { return 256; }
int tc_forward(const nvp_desc*, const LevelTab&, const nvp_params*, const float*, const float*, int64_t, float*, void*,
               size_t, cudaStream_t) { set_error("NVP_MODE_TC_F16 not built"); return 9; }
int tc_fwd_bwd(const nvp_desc*, const LevelTab&, const nvp_params*, const float*, const float*, const uint8_t*,
               const float*, int64_t, int64_t, const nvp_grads*, float*, float*, void*, size_t, cudaStream_t) {
  set_error("NVP_MODE_TC_F16 not built"); return 9; }
}  // namespace nvp
