// fp32 CUDA-core path of the modulated SIREN (mode NVP_MODE_FP32_SIMT).
//
// Keeps the reference's fp32 semantics layer by layer (modulation.py:83-92,112-121; modules.py:81;
// loss_functions.py:3; training.py:47-48,74) with hand-written tiled FFMA GEMMs.  It is the
// exact-arithmetic mode of the library and the on-device cross-check for the tcgen05 path
// (mlp_tc.cu).  Activations live in a caller-provided workspace, processed in sample chunks.
#include <algorithm>

#include "common.cuh"

namespace nvp {
namespace {

constexpr int H = kHidden;
constexpr int64_t kChunk = 131072;  // samples per workspace chunk

// ------------------------------------------------------------------------------------------
// Tiled SGEMM: C[M,N] (+)= A[M,K] * B[K,N], arbitrary strides, 128x128x8 tiles, 8x8 per thread.
// ------------------------------------------------------------------------------------------
struct GemmArgs {
  int M, N, K;
  const float* A; int64_t sam, sak;   // A(m,k) = A[m*sam + k*sak]
  const float* B; int64_t sbk, sbn;   // B(k,n) = B[k*sbk + n*sbn]
  float* C; int64_t ldc;              // C(m,n) = C[m*ldc + n]
  const float* bias;                  // per-n bias or nullptr
  int act;                            // 0 none, 1 LeakyReLU(0.01)
  int accumulate;                     // C = epi(C + A*B)
  int klen;                           // K range per blockIdx.z
  int atomic_out;                     // C += A*B with atomics (split-K / gradient accumulation)
};

constexpr int BM = 128, BN = 128, BK = 8, GT = 256;

template <bool A_MCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(GT) sgemm_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * g.klen;
  const int kend = min(g.K, kbeg + g.klen);
  const int ty = t >> 4, tx = t & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    if constexpr (A_MCONTIG) {
      const int kk = t >> 5, mm = (t & 31) * 4;
      const int k = k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + mm + i;
        As[kk][mm + i] = (k < kend && m < g.M) ? __ldg(g.A + m * g.sam + k * g.sak) : 0.0f;
      }
    } else {
      const int mm = t >> 1, kk = (t & 1) * 4;
      const int m = m0 + mm;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + kk + i;
        As[kk + i][mm] = (k < kend && m < g.M) ? __ldg(g.A + m * g.sam + k * g.sak) : 0.0f;
      }
    }
    if constexpr (B_NCONTIG) {
      const int kk = t >> 5, nn = (t & 31) * 4;
      const int k = k0 + kk;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = n0 + nn + i;
        Bs[kk][nn + i] = (k < kend && n < g.N) ? __ldg(g.B + k * g.sbk + n * g.sbn) : 0.0f;
      }
    } else {
      const int nn = t >> 1, kk = (t & 1) * 4;
      const int n = n0 + nn;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + kk + i;
        Bs[kk + i][nn] = (k < kend && n < g.N) ? __ldg(g.B + k * g.sbk + n * g.sbn) : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool split = g.atomic_out != 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 8 + j;
      if (n >= g.N) continue;
      float* c = g.C + m * g.ldc + n;
      if (split) {
        atomicAdd(c, acc[i][j]);
      } else {
        float v = acc[i][j];
        if (g.accumulate) v += *c;
        if (g.bias) v += __ldg(g.bias + n);
        if (g.act == 1) v = v > 0.0f ? v : 0.01f * v;
        *c = v;
      }
    }
  }
}

int gemm(const GemmArgs& g, int splits, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
  GemmArgs a = g;
  a.klen = (g.K + splits - 1) / splits;
  a.klen = (a.klen + BK - 1) / BK * BK;
  splits = (g.K + a.klen - 1) / a.klen;
  dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN, splits);
  const bool am = (g.sam == 1 && g.sak != 1), bn = (g.sbn == 1);
  ScopedKernelTimer timer(K_SIMT, st);
  if (am && bn) sgemm_kernel<true, true><<<grid, GT, 0, st>>>(a);
  else if (am && !bn) sgemm_kernel<true, false><<<grid, GT, 0, st>>>(a);
  else if (!am && bn) sgemm_kernel<false, true><<<grid, GT, 0, st>>>(a);
  else sgemm_kernel<false, false><<<grid, GT, 0, st>>>(a);
  NVP_LAUNCH_CHECK();
  return 0;
}

// Y[n,H] = act(X[n,K] W[H,K]^T (+ Y) + b)
int linear_nt(const float* X, int ldx, int K, const float* W, int ldw, const float* b, int act, int accumulate,
              float* Y, int64_t n, cudaStream_t st) {
  GemmArgs g{};
  g.M = static_cast<int>(n); g.N = H; g.K = K;
  g.A = X; g.sam = ldx; g.sak = 1;
  g.B = W; g.sbk = 1; g.sbn = ldw;
  g.C = Y; g.ldc = H; g.bias = b; g.act = act; g.accumulate = accumulate;
  return gemm(g, 1, st);
}
// dX[n,K] (+)= dY[n,H] W[H,K]   (W row pitch ldw, column offset folded into W pointer)
int linear_dgrad(const float* dY, const float* W, int ldw, int K, int accumulate, float* dX, int lddx, int64_t n,
                 cudaStream_t st) {
  GemmArgs g{};
  g.M = static_cast<int>(n); g.N = K; g.K = H;
  g.A = dY; g.sam = H; g.sak = 1;
  g.B = W; g.sbk = ldw; g.sbn = 1;
  g.C = dX; g.ldc = lddx; g.accumulate = accumulate;
  return gemm(g, 1, st);
}
// dW[H,K] += dY[n,H]^T X[n,K]   (split over samples, atomics)
int linear_wgrad(const float* dY, const float* X, int ldx, int K, float* dW, int lddw, int64_t n, cudaStream_t st) {
  if (dW == nullptr) return 0;
  GemmArgs g{};
  g.M = H; g.N = K; g.K = static_cast<int>(n);
  g.A = dY; g.sam = 1; g.sak = H;
  g.B = X; g.sbk = ldx; g.sbn = 1;
  g.C = dW; g.ldc = lddw; g.atomic_out = 1;
  const int splits = static_cast<int>(std::max<int64_t>(2, std::min<int64_t>(592, n / 512)));
  return gemm(g, splits, st);
}

// ------------------------------------------------------------------------------------------
// Elementwise / reduction kernels.  Rows are samples, 128 columns, one thread per column.
// ------------------------------------------------------------------------------------------
constexpr int kRowsPerBlock = 64;

// a0 = sin(w0 * (ws0*tau + bs0)) * h0                       (modulation.py:53-56,86-90)
__global__ void __launch_bounds__(H) siren0_fwd_kernel(const float* __restrict__ tau, const float* __restrict__ w,
                                                       const float* __restrict__ b, float w0,
                                                       const float* __restrict__ h0, float* __restrict__ a0, int64_t n) {
  const int j = threadIdx.x;
  const float wj = __ldg(w + j), bj = __ldg(b + j);
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kRowsPerBlock;
  const int64_t r1 = min(n, r0 + kRowsPerBlock);
  for (int64_t r = r0; r < r1; ++r) {
    const float pre = fmaf(__ldg(tau + r), wj, bj);
    a0[r * H + j] = sinf(w0 * pre) * h0[r * H + j];
  }
}

// a = sin(sp) * h
__global__ void __launch_bounds__(256) sin_gate_fwd_kernel(const float* __restrict__ sp, const float* __restrict__ h,
                                                           float* __restrict__ a, int64_t total) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) a[i] = sinf(sp[i]) * h[i];
}

// rgb = a2 Wl^T + bl : one warp per sample.
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ a2, const float* __restrict__ wl,
                                                       const float* __restrict__ bl, float* __restrict__ rgb, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= n) return;
  const float4 a = *reinterpret_cast<const float4*>(a2 + r * H + lane * 4);
  float o[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(wl + c * H + lane * 4));
    float v = a.x * w.x + a.y * w.y + a.z * w.z + a.w * w.w;
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
    o[c] = v + __ldg(bl + c);
  }
  if (lane < 3) rgb[r * 3 + lane] = o[lane];
}

// drgb = dL/d rgb ; loss partial ; da2 = drgb Wl ; dWl += drgb^T a2 ; dbl += sum drgb
// loss mode: drgb = 2 (rgb - (gt-127.5)/127.5) * inv_count ; explicit mode: drgb = dout.
__global__ void __launch_bounds__(H) head_bwd_kernel(const float* __restrict__ rgb, const uint8_t* __restrict__ gt,
                                                     const float* __restrict__ dout, float inv_count,
                                                     const float* __restrict__ wl, const float* __restrict__ a2,
                                                     float* __restrict__ da2, float* __restrict__ dwl,
                                                     float* __restrict__ dbl, float* __restrict__ loss_sum, int64_t n) {
  const int j = threadIdx.x;
  const float w0 = __ldg(wl + j), w1 = __ldg(wl + H + j), w2 = __ldg(wl + 2 * H + j);
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, ls = 0.f;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kRowsPerBlock;
  const int64_t r1 = min(n, r0 + kRowsPerBlock);
  for (int64_t r = r0; r < r1; ++r) {
    float d0, d1, d2;
    if (dout != nullptr) {
      d0 = __ldg(dout + r * 3); d1 = __ldg(dout + r * 3 + 1); d2 = __ldg(dout + r * 3 + 2);
    } else {
      const float t0 = (static_cast<float>(gt[r * 3]) - 127.5f) / 127.5f;
      const float t1 = (static_cast<float>(gt[r * 3 + 1]) - 127.5f) / 127.5f;
      const float t2 = (static_cast<float>(gt[r * 3 + 2]) - 127.5f) / 127.5f;
      const float e0 = rgb[r * 3] - t0, e1 = rgb[r * 3 + 1] - t1, e2 = rgb[r * 3 + 2] - t2;
      ls += e0 * e0 + e1 * e1 + e2 * e2;
      d0 = 2.0f * e0 * inv_count; d1 = 2.0f * e1 * inv_count; d2 = 2.0f * e2 * inv_count;
    }
    const float a = a2[r * H + j];
    da2[r * H + j] = d0 * w0 + d1 * w1 + d2 * w2;
    g0 = fmaf(d0, a, g0); g1 = fmaf(d1, a, g1); g2 = fmaf(d2, a, g2);
    b0 += d0; b1 += d1; b2 += d2;
  }
  if (dwl != nullptr) {
    atomicAdd(dwl + j, g0); atomicAdd(dwl + H + j, g1); atomicAdd(dwl + 2 * H + j, g2);
  }
  if (j == 0) {
    if (dbl != nullptr) { atomicAdd(dbl, b0); atomicAdd(dbl + 1, b1); atomicAdd(dbl + 2, b2); }
    if (loss_sum != nullptr && dout == nullptr) atomicAdd(loss_sum, ls);
  }
}

// Backward through  a = sin(sp) * h ,  h = LeakyReLU(mpre):
//   dsp = da * h * cos(sp) ; dm = (dh_in + da * sin(sp)) * (h > 0 ? 1 : 0.01)
// Column sums of dsp and dm are the bias gradients.
__global__ void __launch_bounds__(H) layer_bwd_kernel(const float* __restrict__ da, const float* __restrict__ dh_in,
                                                      const float* __restrict__ h, const float* __restrict__ sp,
                                                      float* __restrict__ dsp, float* __restrict__ dm,
                                                      float* __restrict__ dbs, float* __restrict__ dbm, int64_t n) {
  const int j = threadIdx.x;
  float sbs = 0.f, sbm = 0.f;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kRowsPerBlock;
  const int64_t r1 = min(n, r0 + kRowsPerBlock);
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t i = r * H + j;
    float s, c;
    sincosf(sp[i], &s, &c);
    const float d = da[i], hv = h[i];
    const float vsp = d * hv * c;
    float dh = d * s;
    if (dh_in != nullptr) dh += dh_in[i];
    const float vm = dh * (hv > 0.0f ? 1.0f : 0.01f);
    dsp[i] = vsp; dm[i] = vm;
    sbs += vsp; sbm += vm;
  }
  if (dbs != nullptr) atomicAdd(dbs + j, sbs);
  if (dbm != nullptr) atomicAdd(dbm + j, sbm);
}

// First SIREN layer: sp0 = w0*(ws0*tau+bs0) is recomputed; d(ws0*tau+bs0) = da*h*cos(sp0)*w0.
__global__ void __launch_bounds__(H) layer0_bwd_kernel(const float* __restrict__ tau, const float* __restrict__ w,
                                                       const float* __restrict__ b, float w0,
                                                       const float* __restrict__ da, const float* __restrict__ dh_in,
                                                       const float* __restrict__ h, float* __restrict__ dm,
                                                       float* __restrict__ dws, float* __restrict__ dbs,
                                                       float* __restrict__ dbm, int64_t n) {
  const int j = threadIdx.x;
  const float wj = __ldg(w + j), bj = __ldg(b + j);
  float sw = 0.f, sb = 0.f, sbm = 0.f;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kRowsPerBlock;
  const int64_t r1 = min(n, r0 + kRowsPerBlock);
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t i = r * H + j;
    const float tv = __ldg(tau + r);
    float s, c;
    sincosf(w0 * fmaf(tv, wj, bj), &s, &c);
    const float d = da[i], hv = h[i];
    const float dpre = d * hv * c * w0;
    sw = fmaf(dpre, tv, sw); sb += dpre;
    const float vm = (dh_in[i] + d * s) * (hv > 0.0f ? 1.0f : 0.01f);
    dm[i] = vm; sbm += vm;
  }
  if (dws != nullptr) atomicAdd(dws + j, sw);
  if (dbs != nullptr) atomicAdd(dbs + j, sb);
  if (dbm != nullptr) atomicAdd(dbm + j, sbm);
}

inline int row_blocks(int64_t n) { return static_cast<int>((n + kRowsPerBlock - 1) / kRowsPerBlock); }

struct Workspace {
  float *z, *h[3], *sp[3], *a[3], *rgb, *dA, *dH, *dSP, *dM, *dZ;
  int ldz;
};

size_t carve(const nvp_desc* d, int64_t chunk, int what, void* base, Workspace* w) {
  const int ldz = round_up(latent_dim(d), 4);
  size_t off = 0;
  auto take = [&](size_t floats) {
    float* p = base ? reinterpret_cast<float*>(static_cast<char*>(base) + off) : nullptr;
    off += (floats * sizeof(float) + 255) / 256 * 256;
    return p;
  };
  Workspace tmp;
  Workspace& ws = w ? *w : tmp;
  ws.ldz = ldz;
  ws.z = take(chunk * ldz);
  for (int i = 0; i < 3; ++i) ws.h[i] = take(chunk * H);
  for (int i = 0; i < 3; ++i) ws.a[i] = take(chunk * H);
  ws.sp[0] = nullptr;
  ws.sp[1] = take(chunk * H);
  ws.sp[2] = take(chunk * H);
  ws.rgb = take(chunk * 3);
  if (what == 1) {
    ws.dA = take(chunk * H); ws.dH = take(chunk * H); ws.dSP = take(chunk * H); ws.dM = take(chunk * H);
    ws.dZ = take(chunk * ldz);
  }
  return off;
}

int forward_chunk(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords, const float* tau,
                  int64_t n, const Workspace& w, float* rgb_out, cudaStream_t st, bool temporal_interp = false) {
  const int Z = latent_dim(d);
  int rc;
  if ((rc = launch_grid_gather(d, tab, p, coords, n, w.z, w.ldz, nullptr, 0, st, temporal_interp))) return rc;
  // modulator layer 0
  if ((rc = linear_nt(w.z, w.ldz, Z, p->mod_w[0], Z, p->mod_b[0], 1, 0, w.h[0], n, st))) return rc;
  siren0_fwd_kernel<<<row_blocks(n), H, 0, st>>>(tau, p->siren_w[0], p->siren_b[0], d->w0_first, w.h[0], w.a[0], n);
  NVP_LAUNCH_CHECK();
  for (int i = 1; i < 3; ++i) {
    // h_i = lrelu(W_i [h_{i-1}; z] + b_i)   (modulation.py:116-119: hidden part first)
    if ((rc = linear_nt(w.h[i - 1], H, H, p->mod_w[i], H + Z, nullptr, 0, 0, w.h[i], n, st))) return rc;
    if ((rc = linear_nt(w.z, w.ldz, Z, p->mod_w[i] + H, H + Z, p->mod_b[i], 1, 1, w.h[i], n, st))) return rc;
    // a_i = sin(Ws_i a_{i-1} + bs_i) * h_i
    if ((rc = linear_nt(w.a[i - 1], H, H, p->siren_w[i], H, p->siren_b[i], 0, 0, w.sp[i], n, st))) return rc;
    const int64_t total = n * H;
    sin_gate_fwd_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, st>>>(w.sp[i], w.h[i], w.a[i], total);
    NVP_LAUNCH_CHECK();
  }
  head_fwd_kernel<<<static_cast<int>((n * 32 + 255) / 256), 256, 0, st>>>(w.a[2], p->last_w, p->last_b, rgb_out, n);
  NVP_LAUNCH_CHECK();
  return 0;
}

}  // namespace

size_t simt_workspace_bytes(const nvp_desc* d, int64_t n, int what) {
  return carve(d, std::min<int64_t>(std::max<int64_t>(n, 1), kChunk), what, nullptr, nullptr);
}

int simt_forward(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords, const float* tsteps,
                 int64_t n, float* out_rgb, void* ws, size_t ws_bytes, cudaStream_t st, bool temporal_interp) {
  NVP_CHECK(ws_bytes >= simt_workspace_bytes(d, n, 0), "workspace too small (see nvp_workspace_bytes)");
  for (int64_t s0 = 0; s0 < n; s0 += kChunk) {
    const int64_t m = std::min(kChunk, n - s0);
    Workspace w;
    carve(d, std::min(n, kChunk), 0, ws, &w);
    int rc = forward_chunk(d, tab, p, coords + 3 * s0, tsteps + s0, m, w, out_rgb + 3 * s0, st, temporal_interp);
    if (rc) return rc;
  }
  return 0;
}

int simt_fwd_bwd(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords, const float* tsteps,
                 const uint8_t* gt_u8, const float* dout, int64_t n, int64_t n_global, const nvp_grads* g,
                 float* loss_sum, float* out_rgb, void* ws, size_t ws_bytes, cudaStream_t st) {
  NVP_CHECK(ws_bytes >= simt_workspace_bytes(d, n, 1), "workspace too small (see nvp_workspace_bytes)");
  const int Z = latent_dim(d);
  const float inv_count = 1.0f / (3.0f * static_cast<float>(n_global));
  for (int64_t s0 = 0; s0 < n; s0 += kChunk) {
    const int64_t m = std::min(kChunk, n - s0);
    Workspace w;
    carve(d, std::min(n, kChunk), 1, ws, &w);
    const float* tau = tsteps + s0;
    float* rgb = out_rgb ? out_rgb + 3 * s0 : w.rgb;
    int rc = forward_chunk(d, tab, p, coords + 3 * s0, tau, m, w, rgb, st);
    if (rc) return rc;

    head_bwd_kernel<<<row_blocks(m), H, 0, st>>>(rgb, gt_u8 ? gt_u8 + 3 * s0 : nullptr, dout ? dout + 3 * s0 : nullptr,
                                                 inv_count, p->last_w, w.a[2], w.dA, g->last_w, g->last_b, loss_sum, m);
    NVP_LAUNCH_CHECK();
    for (int i = 2; i >= 1; --i) {
      // dA holds da_i ; dH holds the modulator-side dh_i coming from layer i+1 (none for i = 2)
      layer_bwd_kernel<<<row_blocks(m), H, 0, st>>>(w.dA, i == 2 ? nullptr : w.dH, w.h[i], w.sp[i], w.dSP, w.dM,
                                                    g->siren_b[i], g->mod_b[i], m);
      NVP_LAUNCH_CHECK();
      if ((rc = linear_wgrad(w.dSP, w.a[i - 1], H, H, g->siren_w[i], H, m, st))) return rc;
      if ((rc = linear_dgrad(w.dSP, p->siren_w[i], H, H, 0, w.dA, H, m, st))) return rc;  // da_{i-1}
      if ((rc = linear_wgrad(w.dM, w.h[i - 1], H, H, g->mod_w[i], H + Z, m, st))) return rc;
      if ((rc = linear_wgrad(w.dM, w.z, w.ldz, Z, g->mod_w[i] ? g->mod_w[i] + H : nullptr, H + Z, m, st))) return rc;
      if ((rc = linear_dgrad(w.dM, p->mod_w[i], H + Z, H, 0, w.dH, H, m, st))) return rc;  // dh_{i-1}
      if ((rc = linear_dgrad(w.dM, p->mod_w[i] + H, H + Z, Z, i == 2 ? 0 : 1, w.dZ, w.ldz, m, st))) return rc;
    }
    layer0_bwd_kernel<<<row_blocks(m), H, 0, st>>>(tau, p->siren_w[0], p->siren_b[0], d->w0_first, w.dA, w.dH, w.h[0],
                                                   w.dM, g->siren_w[0], g->siren_b[0], g->mod_b[0], m);
    NVP_LAUNCH_CHECK();
    if ((rc = linear_wgrad(w.dM, w.z, w.ldz, Z, g->mod_w[0], Z, m, st))) return rc;
    if ((rc = linear_dgrad(w.dM, p->mod_w[0], Z, Z, 1, w.dZ, w.ldz, m, st))) return rc;
    if ((rc = launch_grid_scatter(d, tab, coords + 3 * s0, m, w.dZ, w.ldz, nullptr, 0, 1.0f, nullptr, g, st))) return rc;
  }
  if (cudaEvent_t ev = take_grid_event()) NVP_CUDA(cudaEventRecord(ev, st));   // nvp_record_grid_grads_event
  return 0;
}

}  // namespace nvp
