// extern "C" entry points of libnvp_b200.so (see include/nvp_b200.h for the contract).
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace nvp {

static thread_local std::string g_error;
static thread_local int g_launches = 0;
static thread_local cudaEvent_t g_grid_event = nullptr;

struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  std::vector<int> ids;      // kernel id of record i; events 2i, 2i+1
  size_t used = 0;
};
static thread_local Profiler g_prof;

void prof_start(int id, cudaStream_t st) {
  if (!g_prof.on) return;
  while (g_prof.pool.size() < g_prof.used + 2) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { g_prof.on = false; return; }
    g_prof.pool.push_back(e);
  }
  g_prof.ids.push_back(id);
  cudaEventRecord(g_prof.pool[g_prof.used], st);
}
void prof_stop(cudaStream_t st) {
  if (!g_prof.on || g_prof.ids.size() * 2 != g_prof.used + 2) return;
  cudaEventRecord(g_prof.pool[g_prof.used + 1], st);
  g_prof.used += 2;
}

void set_error(const std::string& msg) { g_error = msg; }
void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }
cudaEvent_t take_grid_event() {
  cudaEvent_t e = g_grid_event;
  g_grid_event = nullptr;
  return e;
}
// Clears the registered event when a backward call leaves without having consumed it (early return, validation error):
// a stale handle must never be recorded by a later call.
struct GridEventGuard {
  ~GridEventGuard() { g_grid_event = nullptr; }
};

int validate_desc(const nvp_desc* d) {
  NVP_CHECK(d != nullptr, "nvp_desc is NULL");
  NVP_CHECK(d->n_levels >= 1 && d->n_levels <= NVP_MAX_LEVELS, "n_levels must be in [1,32]");
  NVP_CHECK(d->n_features == 1 || d->n_features == 2 || d->n_features == 4 || d->n_features == 8,
            "2d n_features_per_level must be 1, 2, 4 or 8");
  NVP_CHECK(d->sparse_features == 1 || d->sparse_features == 2 || d->sparse_features == 4 || d->sparse_features == 8,
            "3d n_features_per_level must be 1, 2, 4 or 8");
  NVP_CHECK(d->t_resolution >= 1 && d->x_resolution >= 1 && d->y_resolution >= 1, "3d resolutions must be >= 1");
  NVP_CHECK(d->hidden == kHidden, "network.n_neurons must be 128");
  NVP_CHECK(d->n_layers == 3, "network.n_hidden_layers must be 3");
  NVP_CHECK(d->base_resolution >= 1 && d->per_level_scale > 0.0f, "bad base_resolution / per_level_scale");
  return 0;
}

// scale_l = exp2f(l * log2f(pls)) * base - 1 ; res_l = ceil(scale_l) + 1   (tcnn grid.h semantics; the
// resolutions equal eval.py:28-35).  fp32 steps with log2/exp2 evaluated in double and rounded once, the
// same recipe as oracle/nvp_oracle.py::level_table so both sides agree bit-for-bit.
int build_level_table(const nvp_desc* d, LevelTab* tab, int64_t* offsets64) {
  if (int rc = validate_desc(d)) return rc;
  const float log2_pls = static_cast<float>(log2(static_cast<double>(d->per_level_scale)));
  int64_t off = 0;
  for (int l = 0; l < d->n_levels; ++l) {
    const float arg = static_cast<float>(l) * log2_pls;
    const float e = static_cast<float>(exp2(static_cast<double>(arg)));
    const float s = e * static_cast<float>(d->base_resolution) - 1.0f;
    const int res = static_cast<int>(ceilf(s)) + 1;
    NVP_CHECK(off + static_cast<int64_t>(res) * res < (1ll << 31), "keyframe plane has more than 2^31 cells");
    if (tab) { tab->scale[l] = s; tab->res[l] = res; tab->offset[l] = static_cast<int32_t>(off); }
    if (offsets64) offsets64[l] = off;
    off += static_cast<int64_t>(res) * res;
  }
  if (tab) { tab->offset[d->n_levels] = static_cast<int32_t>(off); tab->n_levels = d->n_levels; }
  if (offsets64) offsets64[d->n_levels] = off;
  return 0;
}

static int check_device_ptr(const void* p, const char* name) {
  NVP_CHECK(p != nullptr, std::string(name) + " is NULL");
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) { cudaGetLastError(); set_error(std::string(name) + ": not a CUDA pointer"); return 4; }
  NVP_CHECK(attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged,
            std::string(name) + " must be a device pointer (there is no CPU path)");
  return 0;
}

// The grid kernels read and reduce keyframe / sparse-grid entries with 16-byte vector accesses.
static int check_align16(const void* p, const char* name) {
  NVP_CHECK((reinterpret_cast<uintptr_t>(p) & 15u) == 0, std::string(name) + " must be 16-byte aligned");
  return 0;
}

static int check_grid_grads(const nvp_grads* g) {
  int rc;
  if (g->kf_xy && (rc = check_align16(g->kf_xy, "grads.kf_xy"))) return rc;
  if (g->kf_yt && (rc = check_align16(g->kf_yt, "grads.kf_yt"))) return rc;
  if (g->kf_xt && (rc = check_align16(g->kf_xt, "grads.kf_xt"))) return rc;
  if (g->sparse && (rc = check_align16(g->sparse, "grads.sparse"))) return rc;
  return 0;
}

static int check_params(const nvp_params* p) {
  NVP_CHECK(p != nullptr, "nvp_params is NULL");
  int rc;
  if ((rc = check_device_ptr(p->kf_xy, "params.kf_xy"))) return rc;
  if ((rc = check_device_ptr(p->kf_yt, "params.kf_yt"))) return rc;
  if ((rc = check_device_ptr(p->kf_xt, "params.kf_xt"))) return rc;
  if ((rc = check_device_ptr(p->sparse, "params.sparse"))) return rc;
  if ((rc = check_align16(p->kf_xy, "params.kf_xy"))) return rc;
  if ((rc = check_align16(p->kf_yt, "params.kf_yt"))) return rc;
  if ((rc = check_align16(p->kf_xt, "params.kf_xt"))) return rc;
  if ((rc = check_align16(p->sparse, "params.sparse"))) return rc;
  for (int i = 0; i < 3; ++i) {
    if ((rc = check_device_ptr(p->siren_w[i], "params.siren_w"))) return rc;
    if ((rc = check_device_ptr(p->siren_b[i], "params.siren_b"))) return rc;
    if ((rc = check_device_ptr(p->mod_w[i], "params.mod_w"))) return rc;
    if ((rc = check_device_ptr(p->mod_b[i], "params.mod_b"))) return rc;
  }
  if ((rc = check_device_ptr(p->last_w, "params.last_w"))) return rc;
  if ((rc = check_device_ptr(p->last_b, "params.last_b"))) return rc;
  return 0;
}

}  // namespace nvp

using namespace nvp;

extern "C" {

int nvp_version(void) { return 100; }

const char* nvp_last_error(void) { return g_error.c_str(); }

int nvp_last_launch_count(void) { return g_launches; }

int nvp_grid_bin_plan(const nvp_desc* d, int64_t n, int32_t* tiles_per_axis, int32_t* chunk, int32_t* window_extent,
                      int32_t* window_base, size_t* workspace) {
  LevelTab tab;
  if (int rc = build_level_table(d, &tab, nullptr)) return rc;
  int tb = 0, ch = 0;
  grid_bin_plan_info(d, tab, n, &tb, &ch, window_extent, window_base, workspace);
  if (tiles_per_axis) *tiles_per_axis = tb;
  if (chunk) *chunk = ch;
  return 0;
}

int nvp_debug_timeline_read(uint64_t* out, int32_t n) {
  NVP_CHECK(out != nullptr && n > 0, "out is NULL / n <= 0");
  return tc_timeline_read(reinterpret_cast<unsigned long long*>(out), n);
}

int nvp_record_grid_grads_event(void* cuda_event) {
  g_grid_event = static_cast<cudaEvent_t>(cuda_event);
  return 0;
}

int nvp_latent_dim(const nvp_desc* d) { return d ? latent_dim(d) : -1; }

int nvp_profile_enable(int on) {
  g_prof.on = on != 0;
  g_prof.used = 0;
  g_prof.ids.clear();
  return 0;
}

int nvp_profile_read(int max_kinds, float* total_ms, int* counts) {
  NVP_CHECK(total_ms != nullptr && counts != nullptr, "total_ms / counts is NULL");
  for (int k = 0; k < max_kinds; ++k) { total_ms[k] = 0.f; counts[k] = 0; }
  for (size_t i = 0; i < g_prof.used / 2; ++i) {
    NVP_CUDA(cudaEventSynchronize(g_prof.pool[2 * i + 1]));
    float ms = 0.f;
    NVP_CUDA(cudaEventElapsedTime(&ms, g_prof.pool[2 * i], g_prof.pool[2 * i + 1]));
    const int id = g_prof.ids[i];
    if (id >= 0 && id < max_kinds) { total_ms[id] += ms; counts[id] += 1; }
  }
  g_prof.used = 0;
  g_prof.ids.clear();
  return 0;
}

int nvp_level_table(const nvp_desc* d, float* scales, int32_t* res, int64_t* offsets) {
  LevelTab tab;
  int64_t off64[NVP_MAX_LEVELS + 1];
  if (int rc = build_level_table(d, &tab, off64)) return rc;
  for (int l = 0; l < d->n_levels; ++l) {
    if (scales) scales[l] = tab.scale[l];
    if (res) res[l] = tab.res[l];
    if (offsets) offsets[l] = off64[l];
  }
  if (offsets) offsets[d->n_levels] = off64[d->n_levels];
  return 0;
}

int nvp_workspace_bytes(const nvp_desc* d, int64_t n, int mode, int what, size_t* bytes) {
  if (int rc = validate_desc(d)) return rc;
  NVP_CHECK(bytes != nullptr, "bytes is NULL");
  NVP_CHECK(what == 0 || what == 1, "what must be 0 (forward) or 1 (backward)");
  mode &= 0xff;
  if (mode == NVP_MODE_FP32_SIMT) *bytes = simt_workspace_bytes(d, n, what);
  else if (mode == NVP_MODE_TC_F16) *bytes = tc_workspace_bytes(d, n, what);
  else NVP_CHECK(false, "unknown mode");
  return 0;
}

int nvp_encode_latent(const nvp_desc* d, const nvp_params* p, const float* coords, int64_t n, float* z, void* stream) {
  reset_launch_count();
  LevelTab tab;
  if (int rc = build_level_table(d, &tab, nullptr)) return rc;
  NVP_CHECK(n >= 0, "n must be >= 0");
  if (n == 0) return 0;
  if (int rc = check_device_ptr(coords, "coords")) return rc;
  if (int rc = check_device_ptr(z, "z")) return rc;
  NVP_CHECK(p && p->kf_xy && p->kf_yt && p->kf_xt && p->sparse, "grid parameter pointers are NULL");
  return launch_grid_gather(d, tab, p, coords, n, z, latent_dim(d), nullptr, 0, static_cast<cudaStream_t>(stream));
}

int nvp_scatter_latent(const nvp_desc* d, const float* coords, int64_t n, const float* dz, const nvp_grads* g, void* stream) {
  reset_launch_count();
  LevelTab tab;
  if (int rc = build_level_table(d, &tab, nullptr)) return rc;
  NVP_CHECK(n >= 0, "n must be >= 0");
  NVP_CHECK(g != nullptr, "nvp_grads is NULL");
  if (n == 0) return 0;
  if (int rc = check_device_ptr(coords, "coords")) return rc;
  if (int rc = check_device_ptr(dz, "dz")) return rc;
  return launch_grid_scatter(d, tab, coords, n, dz, latent_dim(d), nullptr, 0, 1.0f, nullptr, g,
                             static_cast<cudaStream_t>(stream));
}

int nvp_forward(const nvp_desc* d, const nvp_params* p, const float* coords, const float* tsteps, int64_t n,
                float* out_rgb, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  reset_launch_count();
  LevelTab tab;
  if (int rc = build_level_table(d, &tab, nullptr)) return rc;
  NVP_CHECK(n >= 0, "n must be >= 0");
  if (n == 0) return 0;
  int rc;
  if ((rc = check_params(p))) return rc;
  if ((rc = check_device_ptr(coords, "coords"))) return rc;
  if ((rc = check_device_ptr(tsteps, "tsteps"))) return rc;
  if ((rc = check_device_ptr(out_rgb, "out_rgb"))) return rc;
  if ((rc = check_device_ptr(workspace, "workspace"))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool interp = (mode & NVP_FLAG_TEMPORAL_INTERP) != 0;
  mode &= 0xff;
  if (mode == NVP_MODE_FP32_SIMT) return simt_forward(d, tab, p, coords, tsteps, n, out_rgb, workspace, workspace_bytes, st, interp);
  if (mode == NVP_MODE_TC_F16) return tc_forward(d, tab, p, coords, tsteps, n, out_rgb, workspace, workspace_bytes, st, interp);
  NVP_CHECK(false, "unknown mode");
}

static int fwd_bwd_common(const nvp_desc* d, const nvp_params* p, const float* coords, const float* tsteps,
                          const uint8_t* gt_u8, const float* dout, int64_t n, int64_t n_global, const nvp_grads* g,
                          float* loss_sum, float* out_rgb, void* workspace, size_t workspace_bytes, int mode,
                          void* stream) {
  reset_launch_count();
  GridEventGuard event_guard;   // whatever happens below, the event registered for this call does not outlive it
  LevelTab tab;
  if (int rc = build_level_table(d, &tab, nullptr)) return rc;
  NVP_CHECK(n >= 0 && n_global >= n, "need 0 <= n <= n_global");
  NVP_CHECK(g != nullptr, "nvp_grads is NULL");
  if (n == 0) return 0;
  int rc;
  if ((rc = check_params(p))) return rc;
  if ((rc = check_grid_grads(g))) return rc;
  if ((rc = check_device_ptr(coords, "coords"))) return rc;
  if ((rc = check_device_ptr(tsteps, "tsteps"))) return rc;
  if (gt_u8 && (rc = check_device_ptr(gt_u8, "gt_u8"))) return rc;
  if (dout && (rc = check_device_ptr(dout, "dout"))) return rc;
  if ((rc = check_device_ptr(workspace, "workspace"))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == NVP_MODE_FP32_SIMT)
    return simt_fwd_bwd(d, tab, p, coords, tsteps, gt_u8, dout, n, n_global, g, loss_sum, out_rgb, workspace,
                        workspace_bytes, st);
  if (mode == NVP_MODE_TC_F16)
    return tc_fwd_bwd(d, tab, p, coords, tsteps, gt_u8, dout, n, n_global, g, loss_sum, out_rgb, workspace,
                      workspace_bytes, st);
  NVP_CHECK(false, "unknown mode");
}

int nvp_backward(const nvp_desc* d, const nvp_params* p, const float* coords, const float* tsteps, const float* dout,
                 int64_t n, const nvp_grads* g, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  NVP_CHECK(dout != nullptr || n == 0, "dout is NULL");
  return fwd_bwd_common(d, p, coords, tsteps, nullptr, dout, n, n, g, nullptr, nullptr, workspace, workspace_bytes, mode,
                        stream);
}

int nvp_fwd_loss_bwd(const nvp_desc* d, const nvp_params* p, const float* coords, const float* tsteps,
                     const uint8_t* gt_u8, int64_t n, int64_t n_global, const nvp_grads* g, float* loss_sum,
                     float* out_rgb, void* workspace, size_t workspace_bytes, int mode, void* stream) {
  NVP_CHECK(gt_u8 != nullptr || n == 0, "gt_u8 is NULL");
  NVP_CHECK(loss_sum != nullptr, "loss_sum is NULL");
  return fwd_bwd_common(d, p, coords, tsteps, gt_u8, nullptr, n, n_global, g, loss_sum, out_rgb, workspace,
                        workspace_bytes, mode, stream);
}

}  // extern "C"
