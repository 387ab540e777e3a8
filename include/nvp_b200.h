/*
 * nvp_b200.h — C ABI of the B200-native NVP per-coordinate hot path.
 *
 * Drop-in boundary.  Each entry point replaces a piece of the reference's
 * Python/tiny-cuda-nn path (citations relative to the reference tree):
 *
 *   nvp_forward        modules.py:51-84   NVP.forward  (= tcnn.Encoding x3 modules.py:65-67,
 *                                         SparseGrid.forward sparsegrid.py:23-72,
 *                                         SirenWrapper/Modulator/SirenNet modulation.py:83-92,112-121,142-145)
 *   nvp_backward       training.py:74     autograd backward of NVP.forward for a given dL/d(model_out)
 *   nvp_fwd_loss_bwd   training.py:47-52,74  gt normalise + NVP.forward + loss_functions.py:3 image_mse + backward
 *   nvp_encode_latent  modules.py:61-78   the positional feature vector alone (3 keyframe planes + sparse grid)
 *   nvp_level_table    eval.py:28-35, compression.py:26-33  the DenseGrid level layout (res, offsets, scales)
 *
 * Conventions
 *   - Plain C types only.  All data pointers are DEVICE pointers owned by the caller (torch); the
 *     library never allocates or frees device memory and keeps no global device state.
 *   - `stream` is a cudaStream_t passed as void*.  Calls only enqueue work; they never synchronise.
 *   - Gradients are ACCUMULATED into the caller's buffers (the caller zeroes them).
 *   - Alignment: the keyframe and sparse-grid parameter and gradient buffers must be 16-byte aligned (the grid kernels
 *     use 16-byte vector loads and reductions on them; checked, error code on violation).  torch allocations are.
 *   - Every function returns 0 on success, nonzero on error; nvp_last_error() returns a message
 *     (thread-local).  Nothing is thrown across the boundary.  There is no CPU path: host pointers
 *     are rejected where the driver can tell.
 *   - Layouts (fp32 unless noted):
 *       coords   [N,3]  (t, x, y) in [0,1]          dataio.py:115
 *       tsteps   [N]    SIREN input (t_idx+0.5)/T    dataio.py:113
 *       gt_u8    [N,3]  uint8 RGB                    dataio.py:108
 *       out_rgb  [N,3]
 *       keyframe params: flat [n_cells*F], level-major, cell = i0 + i1*res (input dim 0 fastest),
 *                        feature-minor, no padding   compression.py:72,77
 *       sparse   [T,X,Y,F]                           sparsegrid.py:13
 *       linear weights [out,in] row-major, biases [out]   (torch nn.Linear / modulation.py:40-41)
 */
#ifndef NVP_B200_H_
#define NVP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVP_MAX_LEVELS 32
#define NVP_MAX_LAYERS 3

/* Model description = the slice of config_nvp_{s,l}.json["nvp"] the path reads. */
typedef struct nvp_desc {
  int32_t n_features;      /* 2d_encoding_*.n_features_per_level (2 = config S, 4 = config L) */
  int32_t n_levels;        /* 2d_encoding_*.n_levels (16) */
  int32_t base_resolution; /* 2d_encoding_*.base_resolution (16) */
  float   per_level_scale; /* 2d_encoding_*.per_level_scale (1.35) */
  int32_t sparse_features; /* 3d_encoding.n_features_per_level */
  int32_t t_resolution;    /* 3d_encoding.t_resolution */
  int32_t x_resolution;
  int32_t y_resolution;
  int32_t hidden;          /* network.n_neurons (must be 128) */
  int32_t n_layers;        /* network.n_hidden_layers (must be 3) */
  float   w0_first;        /* 30.0 (modules.py:36) */
} nvp_desc;

/* Parameter / gradient pointer tables.  nvp_grads uses the same field order; a NULL gradient
 * pointer means "do not compute this gradient". */
typedef struct nvp_params {
  const float* kf_xy;                    /* keyframes_xy.params */
  const float* kf_yt;                    /* keyframes_yt.params */
  const float* kf_xt;                    /* keyframes_xt.params */
  const float* sparse;                   /* sparse_grid.embeddings */
  const float* siren_w[NVP_MAX_LAYERS];  /* net.layers.i.weight  [128,1] / [128,128] */
  const float* siren_b[NVP_MAX_LAYERS];  /* net.layers.i.bias */
  const float* last_w;                   /* net.last_layer.weight [3,128] */
  const float* last_b;                   /* net.last_layer.bias   [3] */
  const float* mod_w[NVP_MAX_LAYERS];    /* wrapper.modulator.layers.i.0.weight [128, Z] / [128,128+Z] */
  const float* mod_b[NVP_MAX_LAYERS];    /* wrapper.modulator.layers.i.0.bias */
} nvp_params;

typedef struct nvp_grads {
  float* kf_xy;
  float* kf_yt;
  float* kf_xt;
  float* sparse;
  float* siren_w[NVP_MAX_LAYERS];
  float* siren_b[NVP_MAX_LAYERS];
  float* last_w;
  float* last_b;
  float* mod_w[NVP_MAX_LAYERS];
  float* mod_b[NVP_MAX_LAYERS];
} nvp_grads;

/* Arithmetic mode of the dense layers. */
enum {
  NVP_MODE_FP32_SIMT = 0, /* fp32 FFMA on CUDA cores: bit-faithful to the reference's fp32 semantics */
  NVP_MODE_TC_F16    = 1  /* tcgen05 tensor cores, fp16 operands / fp32 accumulate in TMEM */
};
/* OR-ed into `mode` of nvp_forward: evaluate the 3-D grid with SparseGrid.forward_inter (sparsegrid.py:76-156,
 * selected by NVP.forward(temporal_interp=True), modules.py:72-73; eval-only, no backward). */
#define NVP_FLAG_TEMPORAL_INTERP 0x100

int nvp_version(void);
const char* nvp_last_error(void);

/* DenseGrid level layout.  scales/res/offsets have room for n_levels (+1 for offsets). */
int nvp_level_table(const nvp_desc* d, float* scales, int32_t* res, int64_t* offsets);

/* Latent width Z = 3*n_levels*n_features + 9*sparse_features (modules.py:42-45). */
int nvp_latent_dim(const nvp_desc* d);

/* Bytes of caller-provided scratch needed by a call on n samples (0 = none).
 * `what`: 0 = nvp_forward, 1 = nvp_backward / nvp_fwd_loss_bwd. */
int nvp_workspace_bytes(const nvp_desc* d, int64_t n, int mode, int what, size_t* bytes);

/* z[N, Z] fp32 = [DG_xy(x,y) | DG_yt(t,y) | DG_xt(t,x) | SG(t,x,y)]. */
int nvp_encode_latent(const nvp_desc* d, const nvp_params* p, const float* coords, int64_t n,
                      float* z, void* stream);

/* Backward of nvp_encode_latent: grid gradients += scatter of dz[N, Z] (fp32).  Only the four grid pointers of `g` are
 * used (NULL = skip).  With nvp_encode_latent this is the autograd pair behind the standalone operators the reference
 * calls: tcnn.Encoding.__call__ (modules.py:65-67) and SparseGrid.forward (sparsegrid.py:23-72). */
int nvp_scatter_latent(const nvp_desc* d, const float* coords, int64_t n, const float* dz, const nvp_grads* g,
                       void* stream);

/* out_rgb[N,3] = NVP.forward. */
int nvp_forward(const nvp_desc* d, const nvp_params* p, const float* coords, const float* tsteps,
                int64_t n, float* out_rgb, void* workspace, size_t workspace_bytes, int mode,
                void* stream);

/* grads += d(sum(dout * NVP.forward))/d(params); forward activations are recomputed. */
int nvp_backward(const nvp_desc* d, const nvp_params* p, const float* coords, const float* tsteps,
                 const float* dout, int64_t n, const nvp_grads* g, void* workspace,
                 size_t workspace_bytes, int mode, void* stream);

/* One training step's device work: rgb = forward; loss_sum[0] += sum((rgb-gt)^2) over this call's
 * samples (divide by 3*n_global on the host for image_mse); grads += d(image_mse)/d(params) with the
 * mean taken over 3*n_global elements (n_global >= n: this call may be one shard of the batch).
 * out_rgb may be NULL. */
int nvp_fwd_loss_bwd(const nvp_desc* d, const nvp_params* p, const float* coords, const float* tsteps,
                     const uint8_t* gt_u8, int64_t n, int64_t n_global, const nvp_grads* g,
                     float* loss_sum, float* out_rgb, void* workspace, size_t workspace_bytes,
                     int mode, void* stream);

/* Host-only query of the tile-binned grid plan (nvp_b200/csrc/grid_binned.cuh) the tensor-core mode uses for n samples;
 * no reference counterpart (introspection for tests / tuning).  tiles_per_axis = TB; window_extent[l] = cells per axis of
 * level l's shared-memory window (E_l = ceil(scale_l / TB) + 2, clipped to res_l + 1); window_base[l] = first region
 * cell of level l, window_base[n_levels] = cells per region; chunk = samples per task; workspace = bytes of bucket state.
 * Returns 0 with *tiles_per_axis = 0 when the configuration uses the direct kernels instead.  Any out pointer may be NULL. */
int nvp_grid_bin_plan(const nvp_desc* d, int64_t n, int32_t* tiles_per_axis, int32_t* chunk, int32_t* window_extent,
                      int32_t* window_base, size_t* workspace);

/* Development aid, no reference counterpart: clock64 stamps of the fused forward kernel's phases (CTA 0, 4th tile), only in
 * libraries built with `make TIMELINE=1`; otherwise returns nonzero.  out has room for n (<= 128) host uint64 slots; call
 * after synchronising the stream. */
int nvp_debug_timeline_read(uint64_t* out, int32_t n);

/* Multi-GPU overlap hook (reference counterpart: none - the reference is single-GPU, training.py:74).
 * Inside nvp_backward / nvp_fwd_loss_bwd the grid scatter-add runs before the weight-gradient kernel.  If an event was
 * registered with this call (cudaEvent_t as void*, thread-local, consumed by the next backward call of this thread; NULL
 * clears), it is recorded on the call's stream right after the scatter-add: the keyframe / sparse-grid gradients are
 * then final, so the host can start their collective on another stream while the MLP weight gradients are still
 * being computed; that kernel then leaves NVP_COMM_SMS (env, default 16) SMs free for the collective's CTAs
 * (tensor-core mode; the fp32 mode records the event after its last kernel). */
int nvp_record_grid_grads_event(void* cuda_event);

/* Number of kernels the last call on this thread enqueued (for bench.py's gpu_launches). */
int nvp_last_launch_count(void);

/* Fused AdamW step over flat fp32 buffers (reference training.py:13-14,73-76:
 * torch.optim.AdamW(weight_decay=1e-3) + optim.zero_grad()).  `step` is the 1-based step count used for the bias
 * corrections, `lr` the (cosine-annealed) learning rate of this step; zero_grad != 0 clears `grads` in the same pass.
 * Buffers must be 16-byte aligned. */
int nvp_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int64_t step, int zero_grad,
                   void* stream);

/* Device-resident sampler (reference dataio.py:85-99,104-120 + the per-step host gather and
 * H2D copy of training.py:45-48).  video: uint8 [T, H*W, 3] in device memory; temporal_coords / temporal_steps: the
 * two [T] look-up tables of dataio.py:96,99.  Indices come either from t_idx / p_idx (int64 device arrays, e.g. the
 * reference's CPU RNG stream uploaded: identical batches) or, when both are NULL, from Philox4x32-10 keyed by
 * (seed, step, sample) with frames restricted to [t_lo, t_hi).  Outputs: coords [n,3], tsteps [n], gt [n,3] uint8 and
 * optionally the drawn indices (int32). */
int nvp_sample_batch(const uint8_t* video, int T, int H, int W, const float* temporal_coords,
                     const float* temporal_steps, int64_t n, const int64_t* t_idx, const int64_t* p_idx,
                     uint64_t seed, uint64_t step, int t_lo, int t_hi, float* coords, float* tsteps, uint8_t* gt,
                     int32_t* t_idx_out, int32_t* p_idx_out, void* stream);

/* Per-kernel device timing with CUDA events on the launch stream (bench.py's roofline leg).
 * After nvp_profile_enable(1) every kernel the library enqueues on this thread is bracketed by an
 * event pair; nvp_profile_read synchronises those events, returns per-kind totals and resets.
 * Kinds: 0 weight pack, 1 grid gather, 2 MLP forward, 3 MLP backward (dgrad), 4 MLP wgrad,
 *        5 grid scatter, 6 fp32-mode kernels, 7 misc, 8 sample bucketing, 9 fused MLP forward+loss+backward.
 * No reference counterpart. */
#define NVP_PROFILE_KINDS 10
int nvp_profile_enable(int on);
int nvp_profile_read(int max_kinds, float* total_ms, int* counts);

/* Hardware self-test of the tcgen05 building blocks (one 128x128 tile, fp16 operands, fp32 result D[128,128]):
 *   mode 0: D = A[128,K] * B[128,K]^T  (K-major operands,  K in {64,128,192,256})
 *   mode 1: D = A[K,128]^T * B[K,128]  (MN-major operands, K % 16 == 0, K <= 128)
 *   mode 2+x (x = 0..3): D[:, 0:16] = A[K,128]^T * B[K, 16x:16x+16]  (N = 16 block inside the swizzle atom)
 * No reference counterpart; used by the GPU tests only. */
int nvp_selftest_umma(const void* A, const void* B, float* D, int K, int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NVP_B200_H_ */
