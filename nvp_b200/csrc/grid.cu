// Learnable positional feature lookup: gather (forward) and scatter-add (backward).
//
//   keyframe planes  = tcnn DenseGrid, 2-D, L levels, bilinear   (reference: modules.py:14-23,65-67;
//                      layout eval.py:28-35 / compression.py:72,77; arithmetic SURVEY.md A.2)
//   sparse 3-D grid  = nearest voxel + 3x3 (x,y) neighbourhood   (reference: sparsegrid.py:43-72)
//   latent column order: xy | yt | xt | sparse                   (reference: modules.py:69,78)
//
// This file holds the DIRECT kernels (one scattered access per corner; used by the fp32 mode, the standalone encode /
// scatter entry points and as the fallback of the tensor-core mode) and the host side of the tile-binned kernels of
// grid_binned.cuh, which the tensor-core mode uses for both reference configurations.
//
// Thread mapping of the direct kernels: one thread per (sample, level); the L threads of one sample are adjacent lanes, so
// every plane's L*F output floats are written as one contiguous run and the three coordinate
// loads are warp-broadcasts.  Threads with level index < 9 also fetch one voxel of the 3x3
// neighbourhood.  All table reads go through the read-only path as F-wide vectors; the backward
// uses F-wide vector reductions (red.global.add.v2/v4.f32, sm_90+).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace nvp {
namespace {

template <int F>
__device__ __forceinline__ void ld_feat(const float* __restrict__ p, float (&v)[F]) {
  if constexpr (F == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else if constexpr (F == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (F == 8) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    float4 u = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; v[4] = u.x; v[5] = u.y; v[6] = u.z; v[7] = u.w;
  } else {
#pragma unroll
    for (int f = 0; f < F; ++f) v[f] = __ldg(p + f);
  }
}

template <int F>
__device__ __forceinline__ void red_feat(float* p, const float (&v)[F]) {
  if constexpr (F == 2) {
    atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
  } else if constexpr (F == 4) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  } else if constexpr (F == 8) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    atomicAdd(reinterpret_cast<float4*>(p) + 1, make_float4(v[4], v[5], v[6], v[7]));
  } else {
#pragma unroll
    for (int f = 0; f < F; ++f) atomicAdd(p + f, v[f]);
  }
}

// pos = fma(scale, u, 0.5); cell = floor(pos); frac = pos - cell     
__device__ __forceinline__ void pos_fract(float scale, float u, int& cell, float& frac) {
  float pos = fmaf(scale, u, 0.5f);
  float fl = floorf(pos);
  cell = static_cast<int>(fl);
  frac = pos - fl;
}

// Flat cell index without clamping, wrapped modulo the level size (the tcnn edge aliasing).
__device__ __forceinline__ int wrap_cell(int flat, int cells) {
  if (flat >= cells) flat -= cells;
  if (static_cast<unsigned>(flat) >= static_cast<unsigned>(cells)) {  // inputs outside [0,1]
    flat %= cells;
    if (flat < 0) flat += cells;
  }
  return flat;
}

// nearest voxel: clamp(trunc((res-1)*c + 0.5), 0, res-1), mul and add NOT fused (sparsegrid.py:43-56)
__device__ __forceinline__ int nearest_voxel(float c, int res) {
  float f = __fadd_rn(__fmul_rn(static_cast<float>(res - 1), c), 0.5f);
  int i = __float2int_rz(f);
  return min(max(i, 0), res - 1);
}

struct GridArgs {
  LevelTab tab;
  const float* coords;
  int64_t n;
  const float* kf[3];  // xy, yt, xt (latent order)
  const float* sparse;
  float* gkf[3];
  float* gsparse;
  float* z;            // fp32 latent (gather out) / dz (scatter in)
  int ldz;
  uint8_t* z16t;       // fp16 latent in MMA tile format: [tile][panel][128 rows x 64 halfs, 128B-swizzled]
  int kz;              // panels per tile (ZP / 64)
  int64_t n_pad;       // rows to write (multiple of 128; rows >= n are zero-filled)
  int tres, xres, yres;
  int interp;              // gather: SparseGrid.forward_inter (eval-time temporal blend) instead of nearest frame
  float scale;
  const float* scale_ptr;  // optional device scalar multiplied into scale
  const uint8_t* dz16t;    // scatter input in fp16 MMA tile format (kz panels per 128-sample tile) when non-null
  int lvl_begin;           // scatter: keyframe levels [lvl_begin, L) go straight to global memory
  int n_coarse;            // coarse kernels: levels [0, n_coarse) live in shared memory
  int n_coarse_chunks;     // gather v2: 16-byte chunks per plane handled by the coarse kernel
  int coarse_cells;        // offset[n_coarse]
  int64_t chunk;           // coarse kernel: samples per CTA
  int zero_planes;         // gather_fine last pass: padding rows also clear the keyframe columns (binned path)
};

// Frame selection of the 3-D grid: nearest frame (sparsegrid.py:43-46) or, for temporal_interp, the blend of the
// frames below/above with the reference's coefficient quirk (sparsegrid.py:102-109): upper is normalised first and
// lower is then divided by (new upper + lower); 0/0 = NaN at the last frame, as in the reference.
struct SparseTime { int t0, t1; float w0, w1; };
__device__ __forceinline__ SparseTime sparse_time(const GridArgs& a, float t) {
  SparseTime s;
  if (!a.interp) {
    s.t0 = s.t1 = nearest_voxel(t, a.tres); s.w0 = 1.0f; s.w1 = 0.0f;
    return s;
  }
  const float tf = __fmul_rn(static_cast<float>(a.tres - 1), t);
  const int lower = __float2int_rz(tf);
  const int upper = min(max(__float2int_rz(__fadd_rn(tf, 1.0f)), 0), a.tres - 1);
  float up = __fsub_rn(tf, static_cast<float>(lower)), lo = __fsub_rn(static_cast<float>(upper), tf);
  up = __fdiv_rn(up, __fadd_rn(up, lo));
  lo = __fdiv_rn(lo, __fadd_rn(up, lo));
  s.t0 = min(max(lower, 0), a.tres - 1); s.t1 = upper; s.w0 = lo; s.w1 = up;
  return s;
}
template <int F3>
__device__ __forceinline__ void sparse_fetch(const GridArgs& a, const SparseTime& st, int cx, int cy, float (&fv)[F3]) {
  const size_t v0 = (static_cast<size_t>(st.t0) * a.xres + cx) * a.yres + cy;
  ld_feat<F3>(a.sparse + v0 * F3, fv);
  if (a.interp) {
    const size_t v1 = (static_cast<size_t>(st.t1) * a.xres + cx) * a.yres + cy;
    float fu[F3];
    ld_feat<F3>(a.sparse + v1 * F3, fu);
#pragma unroll
    for (int f = 0; f < F3; ++f) fv[f] = __fadd_rn(__fmul_rn(fv[f], st.w0), __fmul_rn(fu[f], st.w1));
  }
}

struct SampleGeom {
  int it, ix, iy;     // keyframe cell per axis at this level
  float wt, wx, wy;
};

template <int F2>
__device__ __forceinline__ void plane_gather(const float* __restrict__ tab, int off, int res, int i0, float w0,
                                             int i1, float w1, float (&acc)[F2]) {
  const int cells = res * res;
  const float* base = tab + static_cast<size_t>(off) * F2;
  const int b00 = i0 + i1 * res;
  const int c00 = wrap_cell(b00, cells), c10 = wrap_cell(b00 + 1, cells);
  const int c01 = wrap_cell(b00 + res, cells), c11 = wrap_cell(b00 + res + 1, cells);
  float v00[F2], v10[F2], v01[F2], v11[F2];
  ld_feat<F2>(base + static_cast<size_t>(c00) * F2, v00);
  ld_feat<F2>(base + static_cast<size_t>(c10) * F2, v10);
  ld_feat<F2>(base + static_cast<size_t>(c01) * F2, v01);
  ld_feat<F2>(base + static_cast<size_t>(c11) * F2, v11);
  const float a0 = 1.0f - w0, a1 = 1.0f - w1;
  const float k00 = a0 * a1, k10 = w0 * a1, k01 = a0 * w1, k11 = w0 * w1;
#pragma unroll
  for (int f = 0; f < F2; ++f) {
    float r = k00 * v00[f];
    r = fmaf(k10, v10[f], r);
    r = fmaf(k01, v01[f], r);
    r = fmaf(k11, v11[f], r);
    acc[f] = r;
  }
}

template <int F2>
__device__ __forceinline__ void plane_scatter(float* __restrict__ gtab, int off, int res, int i0, float w0, int i1,
                                              float w1, const float (&d)[F2]) {
  const int cells = res * res;
  float* base = gtab + static_cast<size_t>(off) * F2;
  const int b00 = i0 + i1 * res;
  const int c00 = wrap_cell(b00, cells), c10 = wrap_cell(b00 + 1, cells);
  const int c01 = wrap_cell(b00 + res, cells), c11 = wrap_cell(b00 + res + 1, cells);
  const float a0 = 1.0f - w0, a1 = 1.0f - w1;
  const float k00 = a0 * a1, k10 = w0 * a1, k01 = a0 * w1, k11 = w0 * w1;
  if constexpr (F2 == 2) {
    // The two corners of a row are neighbouring cells (16 contiguous bytes): when that pair is 16-byte aligned
    // one red.global.add.v4.f32 replaces two v2 reductions (the L2 reduction rate is the bound of this kernel).
    float* p0 = base + static_cast<size_t>(c00) * 2;
    if (c10 == c00 + 1 && ((off + c00) & 1) == 0) {
      atomicAdd(reinterpret_cast<float4*>(p0), make_float4(k00 * d[0], k00 * d[1], k10 * d[0], k10 * d[1]));
    } else {
      atomicAdd(reinterpret_cast<float2*>(p0), make_float2(k00 * d[0], k00 * d[1]));
      atomicAdd(reinterpret_cast<float2*>(base + static_cast<size_t>(c10) * 2), make_float2(k10 * d[0], k10 * d[1]));
    }
    float* p1 = base + static_cast<size_t>(c01) * 2;
    if (c11 == c01 + 1 && ((off + c01) & 1) == 0) {
      atomicAdd(reinterpret_cast<float4*>(p1), make_float4(k01 * d[0], k01 * d[1], k11 * d[0], k11 * d[1]));
    } else {
      atomicAdd(reinterpret_cast<float2*>(p1), make_float2(k01 * d[0], k01 * d[1]));
      atomicAdd(reinterpret_cast<float2*>(base + static_cast<size_t>(c11) * 2), make_float2(k11 * d[0], k11 * d[1]));
    }
  } else {
    float v[F2];
#pragma unroll
    for (int f = 0; f < F2; ++f) v[f] = k00 * d[f];
    red_feat<F2>(base + static_cast<size_t>(c00) * F2, v);
#pragma unroll
    for (int f = 0; f < F2; ++f) v[f] = k10 * d[f];
    red_feat<F2>(base + static_cast<size_t>(c10) * F2, v);
#pragma unroll
    for (int f = 0; f < F2; ++f) v[f] = k01 * d[f];
    red_feat<F2>(base + static_cast<size_t>(c01) * F2, v);
#pragma unroll
    for (int f = 0; f < F2; ++f) v[f] = k11 * d[f];
    red_feat<F2>(base + static_cast<size_t>(c11) * F2, v);
  }
}

// F consecutive columns [col, col+F) of sample s's latent gradient (fp32 row-major or fp16 tile format).
template <int F>
__device__ __forceinline__ void load_dz(const GridArgs& a, int64_t s, int col, float scale, float (&d)[F]) {
  if (a.dz16t != nullptr) {
    const int64_t tile = s >> 7;
    const int r = static_cast<int>(s & 127);
    const uint8_t* tb = a.dz16t + tile * a.kz * tc::kPanelBytes;
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const int c = col + f;
      d[f] = __half2float(*reinterpret_cast<const __half*>(tb + (c >> 6) * tc::kPanelBytes + tc::panel_offset(r, c & 63))) * scale;
    }
  } else {
    const float* p = a.z + s * a.ldz + col;
#pragma unroll
    for (int f = 0; f < F; ++f) d[f] = __ldg(p + f) * scale;
  }
}

constexpr int kGridThreads = 256;

template <int F2, int F3>
__global__ void __launch_bounds__(kGridThreads) grid_gather_kernel(const GridArgs a) {
  __shared__ float s_scale[NVP_MAX_LEVELS];
  __shared__ int s_res[NVP_MAX_LEVELS];
  __shared__ int s_off[NVP_MAX_LEVELS];
  const int L = a.tab.n_levels;
  if (threadIdx.x < L) {
    s_scale[threadIdx.x] = a.tab.scale[threadIdx.x];
    s_res[threadIdx.x] = a.tab.res[threadIdx.x];
    s_off[threadIdx.x] = a.tab.offset[threadIdx.x];
  }
  __syncthreads();
  const int spb = kGridThreads / L;
  const int ls = threadIdx.x / L, l = threadIdx.x - ls * L;
  const int64_t s = static_cast<int64_t>(blockIdx.x) * spb + ls;
  if (ls >= spb) return;
  if (s >= a.n) {
    // padding rows of the last 128-sample tile: all-zero latent (keeps the wgrad contraction clean)
    if (a.z16t != nullptr && s < a.n_pad) {
      const int64_t tile = s >> 7;
      const int r = static_cast<int>(s & 127);
      for (int c = l * 8; c < a.kz * 64; c += L * 8) {
        uint8_t* dst = a.z16t + (tile * a.kz + (c >> 6)) * tc::kPanelBytes + tc::panel_offset(r, c & 63);
        *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    return;
  }

  const float t = __ldg(a.coords + 3 * s), x = __ldg(a.coords + 3 * s + 1), y = __ldg(a.coords + 3 * s + 2);
  const float sc = s_scale[l];
  const int res = s_res[l], off = s_off[l];
  int it, ix, iy;
  float wt, wx, wy;
  pos_fract(sc, t, it, wt);
  pos_fract(sc, x, ix, wx);
  pos_fract(sc, y, iy, wy);

  float fxy[F2], fyt[F2], fxt[F2];
  plane_gather<F2>(a.kf[0], off, res, ix, wx, iy, wy, fxy);  // xy plane: input (x, y)   modules.py:61
  plane_gather<F2>(a.kf[1], off, res, it, wt, iy, wy, fyt);  // yt plane: input (t, y)   modules.py:63
  plane_gather<F2>(a.kf[2], off, res, it, wt, ix, wx, fxt);  // xt plane: input (t, x)   modules.py:62

  const int pw = L * F2;  // latent columns per plane
  if (a.z != nullptr) {
    float* zr = a.z + s * a.ldz + l * F2;
#pragma unroll
    for (int f = 0; f < F2; ++f) {
      zr[f] = fxy[f];
      zr[pw + f] = fyt[f];
      zr[2 * pw + f] = fxt[f];
    }
  }
  if (a.z16t != nullptr) {
    const int64_t tile = s >> 7;
    const int r = static_cast<int>(s & 127);
    uint8_t* tbase = a.z16t + tile * a.kz * tc::kPanelBytes;
    auto put = [&](int col, const float (&v)[F2]) {
      __half* dst = reinterpret_cast<__half*>(tbase + (col >> 6) * tc::kPanelBytes + tc::panel_offset(r, col & 63));
#pragma unroll
      for (int f = 0; f < F2; ++f) dst[f] = __float2half_rn(v[f]);
    };
    put(l * F2, fxy);
    put(pw + l * F2, fyt);
    put(2 * pw + l * F2, fxt);
    // padding columns [Z, ZP): columns Z and Z+1 carry the constant 1 (bias-gradient column of the wgrad GEMM; the fused
    // kernel's forward weight panels hold the modulator bias there, split hi + lo in fp16), the rest are zero.
    const int zdim = 3 * pw + 9 * F3;
    for (int c = zdim + l; c < a.kz * 64; c += L)
      *reinterpret_cast<__half*>(tbase + (c >> 6) * tc::kPanelBytes + tc::panel_offset(r, c & 63)) =
          __float2half_rn((c == zdim || c == zdim + 1) ? 1.0f : 0.0f);
  }

  // 3x3 neighbourhood of the nearest voxel (same t slice), all weights 1.
  const SparseTime stime = sparse_time(a, t);
  const int vx = nearest_voxel(x, a.xres), vy = nearest_voxel(y, a.yres);
  for (int v = l; v < 9; v += L) {
    const int di = v / 3 - 1, dj = v - (v / 3) * 3 - 1;
    const int cx = min(max(vx + di, 0), a.xres - 1), cy = min(max(vy + dj, 0), a.yres - 1);
    float fv[F3];
    sparse_fetch<F3>(a, stime, cx, cy, fv);
    if (a.z != nullptr) {
#pragma unroll
      for (int f = 0; f < F3; ++f) a.z[s * a.ldz + 3 * pw + v * F3 + f] = fv[f];
    }
    if (a.z16t != nullptr) {
      const int64_t tile = s >> 7;
      const int r = static_cast<int>(s & 127);
      const int col = 3 * pw + v * F3;
      uint8_t* tbase = a.z16t + tile * a.kz * tc::kPanelBytes;
#pragma unroll
      for (int f = 0; f < F3; ++f) {  // per element: col need not be F3-aligned when F2 != F3
        const int c = col + f;
        *reinterpret_cast<__half*>(tbase + (c >> 6) * tc::kPanelBytes + tc::panel_offset(r, c & 63)) =
            __float2half_rn(fv[f]);
      }
    }
  }
}

template <int F2, int F3>
__global__ void __launch_bounds__(kGridThreads) grid_scatter_kernel(const GridArgs a) {
  __shared__ float s_scale[NVP_MAX_LEVELS];
  __shared__ int s_res[NVP_MAX_LEVELS];
  __shared__ int s_off[NVP_MAX_LEVELS];
  const int L = a.tab.n_levels;
  if (threadIdx.x < L) {
    s_scale[threadIdx.x] = a.tab.scale[threadIdx.x];
    s_res[threadIdx.x] = a.tab.res[threadIdx.x];
    s_off[threadIdx.x] = a.tab.offset[threadIdx.x];
  }
  __syncthreads();
  const int spb = kGridThreads / L;
  const int ls = threadIdx.x / L, l = threadIdx.x - ls * L;
  const int64_t s = static_cast<int64_t>(blockIdx.x) * spb + ls;
  if (ls >= spb || s >= a.n) return;

  const float t = __ldg(a.coords + 3 * s), x = __ldg(a.coords + 3 * s + 1), y = __ldg(a.coords + 3 * s + 2);
  const float sc = s_scale[l];
  const int res = s_res[l], off = s_off[l];
  int it, ix, iy;
  float wt, wx, wy;
  pos_fract(sc, t, it, wt);
  pos_fract(sc, x, ix, wx);
  pos_fract(sc, y, iy, wy);

  const int pw = L * F2;
  const float scale = a.scale_ptr ? a.scale * __ldg(a.scale_ptr) : a.scale;
  float d[F2];
  const bool fine = l >= a.lvl_begin;  // coarser levels are handled by grid_scatter_coarse_kernel
  if (a.gkf[0] != nullptr && fine) {
    load_dz<F2>(a, s, l * F2, scale, d);
    plane_scatter<F2>(a.gkf[0], off, res, ix, wx, iy, wy, d);
  }
  if (a.gkf[1] != nullptr && fine) {
    load_dz<F2>(a, s, pw + l * F2, scale, d);
    plane_scatter<F2>(a.gkf[1], off, res, it, wt, iy, wy, d);
  }
  if (a.gkf[2] != nullptr && fine) {
    load_dz<F2>(a, s, 2 * pw + l * F2, scale, d);
    plane_scatter<F2>(a.gkf[2], off, res, it, wt, ix, wx, d);
  }
  if (a.gsparse != nullptr) {
    const int vt = nearest_voxel(t, a.tres), vx = nearest_voxel(x, a.xres), vy = nearest_voxel(y, a.yres);
    for (int v = l; v < 9; v += L) {
      const int di = v / 3 - 1, dj = v - (v / 3) * 3 - 1;
      const int cx = min(max(vx + di, 0), a.xres - 1), cy = min(max(vy + dj, 0), a.yres - 1);
      const size_t vox = (static_cast<size_t>(vt) * a.xres + cx) * a.yres + cy;
      float dv[F3];
      load_dz<F3>(a, s, 3 * pw + v * F3, scale, dv);
      red_feat<F3>(a.gsparse + vox * F3, dv);
    }
  }
}

// Coarse keyframe levels receive millions of reductions into a few hundred cells each (level 0: 256 cells,
// 4 corners x 1.2 M samples x 3 planes); as global reductions they serialise in L2 (measured: level 0 alone
// 1.0 ms of a 2.8 ms scatter).  This kernel keeps a private copy of levels [0, n_coarse) of all three planes
// in shared memory per persistent CTA, accumulates with shared-memory atomics and flushes each CTA's
// non-zero partial sums once.
constexpr int kCoarseThreads = 512;

template <int F2>
__device__ __forceinline__ void plane_scatter_smem(float* __restrict__ tab, int off, int res, int i0, float w0, int i1,
                                                   float w1, const float (&d)[F2]) {
  const int cells = res * res;
  float* base = tab + static_cast<size_t>(off) * F2;
  const int b00 = i0 + i1 * res;
  const int c00 = wrap_cell(b00, cells), c10 = wrap_cell(b00 + 1, cells);
  const int c01 = wrap_cell(b00 + res, cells), c11 = wrap_cell(b00 + res + 1, cells);
  const float a0 = 1.0f - w0, a1 = 1.0f - w1;
  const float k00 = a0 * a1, k10 = w0 * a1, k01 = a0 * w1, k11 = w0 * w1;
#pragma unroll
  for (int f = 0; f < F2; ++f) {
    atomicAdd(base + c00 * F2 + f, k00 * d[f]);
    atomicAdd(base + c10 * F2 + f, k10 * d[f]);
    atomicAdd(base + c01 * F2 + f, k01 * d[f]);
    atomicAdd(base + c11 * F2 + f, k11 * d[f]);
  }
}

template <int F2>
__global__ void __launch_bounds__(kCoarseThreads) grid_scatter_coarse_kernel(const GridArgs a) {
  extern __shared__ float s_tab[];  // [3 planes][coarse_cells][F2]
  __shared__ float s_scale[NVP_MAX_LEVELS];
  __shared__ int s_res[NVP_MAX_LEVELS];
  __shared__ int s_off[NVP_MAX_LEVELS];
  const int LC = a.n_coarse;
  if (threadIdx.x < LC) {
    s_scale[threadIdx.x] = a.tab.scale[threadIdx.x];
    s_res[threadIdx.x] = a.tab.res[threadIdx.x];
    s_off[threadIdx.x] = a.tab.offset[threadIdx.x];
  }
  const int plane_floats = a.coarse_cells * F2;
  for (int i = threadIdx.x; i < 3 * plane_floats; i += kCoarseThreads) s_tab[i] = 0.0f;
  __syncthreads();

  const int pw = a.tab.n_levels * F2;
  const float scale = a.scale_ptr ? a.scale * __ldg(a.scale_ptr) : a.scale;
  const int spp = kCoarseThreads / LC;  // samples per pass
  const int ls = threadIdx.x / LC, l = threadIdx.x - ls * LC;
  const int64_t s_begin = static_cast<int64_t>(blockIdx.x) * a.chunk;
  const int64_t s_end = min(a.n, s_begin + a.chunk);
  if (ls < spp) {
    const float sc = s_scale[l];
    const int res = s_res[l], off = s_off[l];
    for (int64_t s = s_begin + ls; s < s_end; s += spp) {
      const float t = __ldg(a.coords + 3 * s), x = __ldg(a.coords + 3 * s + 1), y = __ldg(a.coords + 3 * s + 2);
      int it, ix, iy;
      float wt, wx, wy;
      pos_fract(sc, t, it, wt);
      pos_fract(sc, x, ix, wx);
      pos_fract(sc, y, iy, wy);
      float d[F2];
      if (a.gkf[0] != nullptr) {
        load_dz<F2>(a, s, l * F2, scale, d);
        plane_scatter_smem<F2>(s_tab, off, res, ix, wx, iy, wy, d);
      }
      if (a.gkf[1] != nullptr) {
        load_dz<F2>(a, s, pw + l * F2, scale, d);
        plane_scatter_smem<F2>(s_tab + plane_floats, off, res, it, wt, iy, wy, d);
      }
      if (a.gkf[2] != nullptr) {
        load_dz<F2>(a, s, 2 * pw + l * F2, scale, d);
        plane_scatter_smem<F2>(s_tab + 2 * plane_floats, off, res, it, wt, ix, wx, d);
      }
    }
  }
  __syncthreads();
  for (int p = 0; p < 3; ++p) {
    if (a.gkf[p] == nullptr) continue;
    for (int i = threadIdx.x; i < plane_floats; i += kCoarseThreads) {
      const float v = s_tab[p * plane_floats + i];
      if (v != 0.0f) atomicAdd(a.gkf[p] + i, v);
    }
  }
}

template <int F2>
int launch_coarse(const GridArgs& a, int blocks, size_t smem, cudaStream_t st) {
  NVP_CUDA(cudaFuncSetAttribute(grid_scatter_coarse_kernel<F2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  grid_scatter_coarse_kernel<F2><<<blocks, kCoarseThreads, smem, st>>>(a);
  return 0;
}

// ------------------------------------------------------------------------------------------
// Tile-format gather, v2 (used when n_levels * F2 is a multiple of 8, i.e. both reference configs).
// A 16-byte chunk of a latent row holds LPC = 8 / F2 consecutive levels of one plane, so one thread produces one
// whole chunk (one st.global.v4) instead of F2 halfs.
//   * gather_coarse_kernel: persistent CTAs, one keyframe plane each; levels [0, LC) of that plane are staged in
//     shared memory (<= 180 KB) and serve the corner reads of the first CC chunks (44 % of all corner reads for S)
//     without touching L2.
//   * gather_fine_kernel: blockIdx.y walks (plane, chunk) pairs for the remaining levels - the tables of one pass
//     (<= 34 MB for S) stay L2-resident while every sample is processed - plus one pass for the 3x3 voxel
//     neighbourhood and the padding columns.
// ------------------------------------------------------------------------------------------
constexpr int kCoarseGatherThreads = 1024;

template <int F2>
__device__ __forceinline__ void bilinear_any(const float* __restrict__ gtab, const float* __restrict__ stab, bool in_smem,
                                             int off, int res, int i0, float w0, int i1, float w1, float (&acc)[F2]) {
  const int cells = res * res;
  const int b00 = i0 + i1 * res;
  const int c00 = wrap_cell(b00, cells), c10 = wrap_cell(b00 + 1, cells);
  const int c01 = wrap_cell(b00 + res, cells), c11 = wrap_cell(b00 + res + 1, cells);
  float v00[F2], v10[F2], v01[F2], v11[F2];
  if (in_smem) {
    const float* base = stab + static_cast<size_t>(off) * F2;
#pragma unroll
    for (int f = 0; f < F2; ++f) {
      v00[f] = base[c00 * F2 + f]; v10[f] = base[c10 * F2 + f]; v01[f] = base[c01 * F2 + f]; v11[f] = base[c11 * F2 + f];
    }
  } else {
    const float* base = gtab + static_cast<size_t>(off) * F2;
    ld_feat<F2>(base + static_cast<size_t>(c00) * F2, v00);
    ld_feat<F2>(base + static_cast<size_t>(c10) * F2, v10);
    ld_feat<F2>(base + static_cast<size_t>(c01) * F2, v01);
    ld_feat<F2>(base + static_cast<size_t>(c11) * F2, v11);
  }
  const float a0 = 1.0f - w0, a1 = 1.0f - w1;
  const float k00 = a0 * a1, k10 = w0 * a1, k01 = a0 * w1, k11 = w0 * w1;
#pragma unroll
  for (int f = 0; f < F2; ++f) {
    float r = k00 * v00[f];
    r = fmaf(k10, v10[f], r);
    r = fmaf(k01, v01[f], r);
    r = fmaf(k11, v11[f], r);
    acc[f] = r;
  }
}

// One 16-byte chunk (levels [chunk*LPC, chunk*LPC+LPC) of `plane`) of sample s's latent row.
template <int F2>
__device__ __forceinline__ void gather_chunk(const GridArgs& a, const float* s_scale, const int* s_res, const int* s_off,
                                             const float* stab, int n_smem_levels, int plane, int chunk, int64_t s) {
  constexpr int LPC = 8 / F2;
  const int64_t tile = s >> 7;
  const int r = static_cast<int>(s & 127);
  const int col = plane * a.tab.n_levels * F2 + chunk * 8;
  uint8_t* dst = a.z16t + (tile * a.kz + (col >> 6)) * tc::kPanelBytes + tc::panel_chunk_offset(r, (col & 63) >> 3);
  if (s >= a.n) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const float t = __ldg(a.coords + 3 * s), x = __ldg(a.coords + 3 * s + 1), y = __ldg(a.coords + 3 * s + 2);
  const float u0 = plane == 0 ? x : t;          // xy: (x, y); yt: (t, y); xt: (t, x)   modules.py:61-63
  const float u1 = plane == 2 ? x : y;
  float out[8];
#pragma unroll
  for (int j = 0; j < LPC; ++j) {
    const int l = chunk * LPC + j;
    int i0, i1;
    float w0, w1;
    pos_fract(s_scale[l], u0, i0, w0);
    pos_fract(s_scale[l], u1, i1, w1);
    float acc[F2];
    bilinear_any<F2>(a.kf[plane], stab, l < n_smem_levels, s_off[l], s_res[l], i0, w0, i1, w1, acc);
#pragma unroll
    for (int f = 0; f < F2; ++f) out[j * F2 + f] = acc[f];
  }
  uint4 q;
  q.x = tc::pack_half2(out[0], out[1]); q.y = tc::pack_half2(out[2], out[3]);
  q.z = tc::pack_half2(out[4], out[5]); q.w = tc::pack_half2(out[6], out[7]);
  *reinterpret_cast<uint4*>(dst) = q;
}

template <int F2>
__global__ void __launch_bounds__(kCoarseGatherThreads) gather_coarse_kernel(const GridArgs a) {
  extern __shared__ float s_tab[];
  __shared__ float s_scale[NVP_MAX_LEVELS];
  __shared__ int s_res[NVP_MAX_LEVELS];
  __shared__ int s_off[NVP_MAX_LEVELS];
  if (threadIdx.x < a.tab.n_levels) {
    s_scale[threadIdx.x] = a.tab.scale[threadIdx.x];
    s_res[threadIdx.x] = a.tab.res[threadIdx.x];
    s_off[threadIdx.x] = a.tab.offset[threadIdx.x];
  }
  const int plane = blockIdx.x % 3;
  const int rank = blockIdx.x / 3, nrank = (gridDim.x - plane + 2) / 3;   // CTAs serving this plane
  const int nfl = a.coarse_cells * F2;
  for (int i = threadIdx.x; i < nfl; i += kCoarseGatherThreads) s_tab[i] = __ldg(a.kf[plane] + i);
  __syncthreads();
  const int64_t per = (a.n_pad + nrank - 1) / nrank;
  const int64_t s0 = rank * per, s1 = min(a.n_pad, s0 + per);
  const int64_t items = (s1 - s0) * a.n_coarse_chunks;
  for (int64_t i = threadIdx.x; i < items; i += kCoarseGatherThreads) {
    const int chunk = static_cast<int>(i / (s1 - s0));
    const int64_t s = s0 + (i - chunk * (s1 - s0));
    gather_chunk<F2>(a, s_scale, s_res, s_off, s_tab, a.n_coarse, plane, chunk, s);
  }
}

// 3x3 neighbourhood of the nearest voxel + padding columns (constant 1 at columns Z and Z+1, then zeros) of one latent row,
// assembled in registers and written as whole 16-byte chunks (c0 = first voxel column = 3*L*F2, chunk aligned).
template <int F3>
__device__ __forceinline__ void sparse_pad_sample(const GridArgs& a, int c0, int64_t s) {
  const int64_t tile = s >> 7;
  const int r = static_cast<int>(s & 127);
  uint8_t* tbase = a.z16t + tile * a.kz * tc::kPanelBytes;
  constexpr int NV = 9 * F3;                 // voxel features
  constexpr int NCH = (NV + 2 + 7) / 8;      // chunks holding features + the two constant-1 columns
  auto chunk_ptr = [&](int c) {
    return reinterpret_cast<uint4*>(tbase + (c >> 6) * tc::kPanelBytes + tc::panel_chunk_offset(r, (c & 63) >> 3));
  };
  const int total_chunks = (a.kz * 64 - c0) >> 3;
  if (s >= a.n) {
    for (int j = 0; j < total_chunks; ++j) *chunk_ptr(c0 + 8 * j) = make_uint4(0u, 0u, 0u, 0u);
    if (a.zero_planes)
      for (int c = 0; c < c0; c += 8) *chunk_ptr(c) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const float t = __ldg(a.coords + 3 * s), x = __ldg(a.coords + 3 * s + 1), y = __ldg(a.coords + 3 * s + 2);
  const SparseTime stime = sparse_time(a, t);
  const int vx = nearest_voxel(x, a.xres), vy = nearest_voxel(y, a.yres);
  float vals[NCH * 8];
#pragma unroll
  for (int i = 0; i < NCH * 8; ++i) vals[i] = (i == NV || i == NV + 1) ? 1.0f : 0.0f;
#pragma unroll
  for (int v = 0; v < 9; ++v) {
    const int di = v / 3 - 1, dj = v - (v / 3) * 3 - 1;
    const int cx = min(max(vx + di, 0), a.xres - 1), cy = min(max(vy + dj, 0), a.yres - 1);
    float fv[F3];
    sparse_fetch<F3>(a, stime, cx, cy, fv);
#pragma unroll
    for (int f = 0; f < F3; ++f) vals[v * F3 + f] = fv[f];
  }
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    if (j < total_chunks) {
      uint4 q;
      q.x = tc::pack_half2(vals[8 * j + 0], vals[8 * j + 1]); q.y = tc::pack_half2(vals[8 * j + 2], vals[8 * j + 3]);
      q.z = tc::pack_half2(vals[8 * j + 4], vals[8 * j + 5]); q.w = tc::pack_half2(vals[8 * j + 6], vals[8 * j + 7]);
      *chunk_ptr(c0 + 8 * j) = q;
    }
  }
  for (int j = NCH; j < total_chunks; ++j) *chunk_ptr(c0 + 8 * j) = make_uint4(0u, 0u, 0u, 0u);
}

template <int F2, int F3>
__global__ void __launch_bounds__(kGridThreads) gather_fine_kernel(const GridArgs a) {
  __shared__ float s_scale[NVP_MAX_LEVELS];
  __shared__ int s_res[NVP_MAX_LEVELS];
  __shared__ int s_off[NVP_MAX_LEVELS];
  const int L = a.tab.n_levels;
  if (threadIdx.x < L) {
    s_scale[threadIdx.x] = a.tab.scale[threadIdx.x];
    s_res[threadIdx.x] = a.tab.res[threadIdx.x];
    s_off[threadIdx.x] = a.tab.offset[threadIdx.x];
  }
  __syncthreads();
  const int64_t s = static_cast<int64_t>(blockIdx.x) * kGridThreads + threadIdx.x;
  if (s >= a.n_pad) return;
  constexpr int LPC = 8 / F2;
  const int cpp = L / LPC;                        // chunks per plane
  const int fine_per_plane = cpp - a.n_coarse_chunks;
  const int pass = blockIdx.y;
  if (pass < 3 * fine_per_plane) {
    // chunk-major so that consecutive passes reuse the same levels' working set size class; plane fastest
    const int chunk = a.n_coarse_chunks + pass / 3, plane = pass % 3;
    gather_chunk<F2>(a, s_scale, s_res, s_off, nullptr, 0, plane, chunk, s);
    return;
  }
  sparse_pad_sample<F3>(a, 3 * L * F2, s);
}

template <int F2>
int launch_gather_v2(const GridArgs& a0, int f3, cudaStream_t st) {
  GridArgs a = a0;
  constexpr int LPC = 8 / F2;
  const int L = a.tab.n_levels;
  int lc = 0;
  while (lc < L && static_cast<size_t>(a.tab.offset[lc + 1]) * F2 * sizeof(float) <= 180 * 1024) ++lc;
  a.n_coarse = lc;
  a.coarse_cells = a.tab.offset[lc];
  a.n_coarse_chunks = std::min(L / LPC, (lc + LPC - 1) / LPC);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (a.n_coarse_chunks > 0) {
    const size_t smem = static_cast<size_t>(a.coarse_cells) * F2 * sizeof(float);
    NVP_CUDA(cudaFuncSetAttribute(gather_coarse_kernel<F2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int blocks = static_cast<int>(std::max<int64_t>(3, std::min<int64_t>(sms, (a.n_pad + 255) / 256 * 3)));
    ScopedKernelTimer timer(K_GATHER, st);
    gather_coarse_kernel<F2><<<blocks, kCoarseGatherThreads, smem, st>>>(a);
    NVP_LAUNCH_CHECK();
  }
  const int passes = 3 * (L / LPC - a.n_coarse_chunks) + 1;
  dim3 grid(static_cast<unsigned>((a.n_pad + kGridThreads - 1) / kGridThreads), passes);
  ScopedKernelTimer timer(K_GATHER, st);
  switch (f3) {
    case 1: gather_fine_kernel<F2, 1><<<grid, kGridThreads, 0, st>>>(a); break;
    case 2: gather_fine_kernel<F2, 2><<<grid, kGridThreads, 0, st>>>(a); break;
    case 4: gather_fine_kernel<F2, 4><<<grid, kGridThreads, 0, st>>>(a); break;
    case 8: gather_fine_kernel<F2, 8><<<grid, kGridThreads, 0, st>>>(a); break;
    default: NVP_CHECK(false, "3d n_features_per_level must be 1, 2, 4 or 8");
  }
  NVP_LAUNCH_CHECK();
  return 0;
}

template <int F2, int F3>
void launch_pair(bool scatter, const GridArgs& a, int blocks, cudaStream_t st) {
  if (scatter)
    grid_scatter_kernel<F2, F3><<<blocks, kGridThreads, 0, st>>>(a);
  else
    grid_gather_kernel<F2, F3><<<blocks, kGridThreads, 0, st>>>(a);
}

template <int F2>
int dispatch_f3(bool scatter, int f3, const GridArgs& a, int blocks, cudaStream_t st) {
  switch (f3) {
    case 1: launch_pair<F2, 1>(scatter, a, blocks, st); return 0;
    case 2: launch_pair<F2, 2>(scatter, a, blocks, st); return 0;
    case 4: launch_pair<F2, 4>(scatter, a, blocks, st); return 0;
    case 8: launch_pair<F2, 8>(scatter, a, blocks, st); return 0;
  }
  return 1;
}

int dispatch(bool scatter, int f2, int f3, const GridArgs& a, cudaStream_t st) {
  const int spb = kGridThreads / a.tab.n_levels;
  const int64_t rows = (!scatter && a.z16t != nullptr) ? a.n_pad : a.n;
  const int blocks = static_cast<int>((rows + spb - 1) / spb);
  if (blocks == 0) return 0;
  int rc = 1;
  ScopedKernelTimer timer(scatter ? K_SCATTER : K_GATHER, st);
  switch (f2) {
    case 1: rc = dispatch_f3<1>(scatter, f3, a, blocks, st); break;
    case 2: rc = dispatch_f3<2>(scatter, f3, a, blocks, st); break;
    case 4: rc = dispatch_f3<4>(scatter, f3, a, blocks, st); break;
    case 8: rc = dispatch_f3<8>(scatter, f3, a, blocks, st); break;
  }
  NVP_CHECK(rc == 0, "n_features_per_level must be 1, 2, 4 or 8");
  NVP_LAUNCH_CHECK();
  return 0;
}

#include "grid_halfs.cuh"

// Voxel-neighbourhood scatter-add of the binned path: one thread per (sample, x-row of the 3x3 neighbourhood).  The three
// voxels of a row are contiguous in memory (y fastest), so for F3 == 2 the row's 6 floats go out as one 16-byte and
// one 8-byte reduction (2 instead of 3 lane-level reductions).  idx = 3 * sample + row.
// Split into the loads (coordinates, the row's three gradient values as raw bits) and the reductions, so that a caller
// can put the loads of several rows in flight before the first reduction (the compiler does not move loads across them).
template <int F3> struct SparseRowIn { float t, x, y; RawHalfs<F3> q[3]; };
template <int F3>
__device__ __forceinline__ SparseRowIn<F3> sparse_row_load(const GridArgs& a, int col0, int64_t idx) {
  const int64_t s = idx / 3;
  const int i = static_cast<int>(idx - s * 3);   // row offset + 1
  SparseRowIn<F3> in;
  in.t = __ldg(a.coords + 3 * s); in.x = __ldg(a.coords + 3 * s + 1); in.y = __ldg(a.coords + 3 * s + 2);
  const uint8_t* tb = a.dz16t + (s >> 7) * a.kz * tc::kPanelBytes;
  const int r = static_cast<int>(s & 127);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int c = col0 + (i * 3 + j) * F3;
    in.q[j] = ld_halfs_raw<F3>(tb + (c >> 6) * tc::kPanelBytes + tc::panel_offset(r, c & 63));
  }
  return in;
}
template <int F3>
__device__ __forceinline__ void sparse_row_commit(const GridArgs& a, float scale, int64_t idx, const SparseRowIn<F3>& in) {
  const int i = static_cast<int>(idx % 3);
  const int vt = nearest_voxel(in.t, a.tres), vx = nearest_voxel(in.x, a.xres), vy = nearest_voxel(in.y, a.yres);
  const int cx = min(max(vx + i - 1, 0), a.xres - 1);
  float d[3][F3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    cvt_halfs<F3>(in.q[j], d[j]);
#pragma unroll
    for (int f = 0; f < F3; ++f) d[j][f] *= scale;
  }
  float* row = a.gsparse + (static_cast<size_t>(vt) * a.xres + cx) * a.yres * F3;
  if (F3 == 2 && vy >= 1 && vy + 1 < a.yres) {
    float* p0 = row + static_cast<size_t>(vy - 1) * 2;
    if ((reinterpret_cast<uintptr_t>(p0) & 15) == 0) {
      atomicAdd(reinterpret_cast<float4*>(p0), make_float4(d[0][0], d[0][1], d[1][0], d[1][1]));
      atomicAdd(reinterpret_cast<float2*>(p0 + 4), make_float2(d[2][0], d[2][1]));
    } else {
      atomicAdd(reinterpret_cast<float2*>(p0), make_float2(d[0][0], d[0][1]));
      atomicAdd(reinterpret_cast<float4*>(p0 + 2), make_float4(d[1][0], d[1][1], d[2][0], d[2][1]));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int cy = min(max(vy + j - 1, 0), a.yres - 1);
      red_feat<F3>(row + static_cast<size_t>(cy) * F3, d[j]);
    }
  }
}
template <int F3>
__device__ __forceinline__ void sparse_scatter_row(const GridArgs& a, int col0, float scale, int64_t idx) {
  sparse_row_commit<F3>(a, scale, idx, sparse_row_load<F3>(a, col0, idx));
}

template <int F3>
__global__ void __launch_bounds__(256) sparse_scatter_rows_kernel(const GridArgs a, int col0) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= 3 * a.n) return;
  const float scale = a.scale_ptr ? a.scale * __ldg(a.scale_ptr) : a.scale;
  sparse_scatter_row<F3>(a, col0, scale, idx);
}

#include "grid_binned.cuh"

// ---- host side of the binned path -----------------------------------------------------------
struct BinPlan {
  BinTab bt;
  int warps;            // window warps (= private regions) per CTA
  int sp_warps;         // extra warps per CTA running the voxel-neighbourhood role (0: separate kernels)
  size_t smem;          // dynamic shared memory per CTA
  int max_tasks;
  size_t o_cnt, o_offs, o_cursor, o_ntasks, o_zeros, o_scan, o_tasks, o_recs, total;   // workspace byte offsets
};

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// false: this configuration does not use the binned path (the direct kernels serve it).
// `scatter` selects the launch shape of the scatter-add kernel (its register footprint is larger); tile size, chunking and
// the workspace layout do not depend on it.
bool plan_bins(const nvp_desc* d, const LevelTab& tab, int64_t n, BinPlan* pl, bool scatter = false) {
  const int L = tab.n_levels, F2 = d->n_features;
  if (env_int("NVP_GRID_BINNED", 1) == 0) return false;
  // 32-bit row offsets into the latent tile buffer (<= 4 panels per 128-sample tile) and 32-bit bucket positions
  if ((L * F2) % 8 != 0 || n < 1 || n > 8000000) return false;
  constexpr size_t kRegionBudget = 216 * 1024;
  const int forced_tb = env_int("NVP_BIN_TB", 0);
  const int want_warps = 20;   // a tile size qualifies when this many private windows fit one SM
  BinTab bt{};
  int warps = 0;
  for (int tb = 32; tb <= 128; tb <<= 1) {   // (3 * tb^2 buckets: a multiple of the scan kernel's 1024 per CTA)
    if (forced_tb && tb != forced_tb) continue;
    int base = 0;
    for (int l = 0; l < L; ++l) {
      int E = static_cast<int>(ceilf(tab.scale[l] / static_cast<float>(tb))) + 2;
      E = std::max(3, std::min(E, tab.res[l] + 1));
      bt.E[l] = E;
      bt.base[l] = base;
      bt.magic[l] = (1u << 20) / static_cast<uint32_t>(E) + 1u;
      for (int idx = 0; idx < E * E; ++idx)   // the reciprocal trick must be exact on the whole window
        if (static_cast<int>((static_cast<uint32_t>(idx) * bt.magic[l]) >> 20) != idx / E) return false;
      base += E * E;
    }
    bt.base[L] = base;
    bt.tb = tb; bt.nt = tb * tb;
    bt.log_tb = 0;
    while ((1 << bt.log_tb) < tb) ++bt.log_tb;
    const size_t region = (static_cast<size_t>((base * F2 + 3) & ~3) + kStageFloats) * sizeof(float);   // + batch stage
    warps = static_cast<int>(kRegionBudget / region);
    if (warps >= want_warps || forced_tb || tb == 128) break;
  }
  if (warps < 1) return false;
  // measured on config S (B200): gather 20 window + 4 voxel warps; scatter-add 16 interleaved windows + 2 (22 + 2 packed)
  int sp_warps = (d->sparse_features == F2 && F2 <= 4) ? env_int(scatter ? "NVP_BIN_SPARSE_WARPS_S" : "NVP_BIN_SPARSE_WARPS_G", scatter ? 2 : 4) : 0;
  sp_warps = std::max(0, std::min(sp_warps, kBinThreadsMax / 32 - 1));
  // (the shared-memory limit of the scatter-add's layout is applied below)
  warps = std::max(1, std::min(std::min(warps, kBinThreadsMax / 32 - sp_warps),
                               env_int(scatter ? "NVP_BIN_WARPS_S" : "NVP_BIN_WARPS_G", scatter ? 22 : 20)));
  size_t region_floats = static_cast<size_t>((bt.base[L] * F2 + 3) & ~3), table_bytes = 0;
  if (scatter) {
    // the scatter-add stages the samples' latent-gradient slices (NC 16-byte chunks each) in a 2 x 32-sample ring per warp
    constexpr size_t kScatterBudget = 226 * 1024;   // all of an SM's shared memory but the kernel's static tables
    const int NC = L * F2 / 8;
    const size_t ring_floats = 2 * 32 * static_cast<size_t>(NC) * 4;
    const size_t idle_floats = (L % 16 != 0) ? 32 * 2 * static_cast<size_t>(F2) : 0;
    bt.dz_extra_floats = static_cast<int32_t>(ring_floats + idle_floats);
    bt.dz_base = static_cast<int32_t>(region_floats + stage_floats(scatter));
    bt.idle_base = static_cast<int32_t>(region_floats + stage_floats(scatter) + ring_floats);
    bt.dz_stride = NC * 16;
    const int cap = warps;
    warps = std::min(cap, static_cast<int>(kScatterBudget / ((region_floats + stage_floats(scatter) + ring_floats + idle_floats) * sizeof(float))));
    if (warps < 1) return false;
    if (F2 == 2 && L == 16 && env_int("NVP_BIN_ILV", 1) != 0) {   // (16 levels: every lane has a level)
      // bank-interleaved windows (grid_binned.cuh): region rows = the largest window of each level group.  The ring lives
      // in the rows of group 1 that the group's first NC levels (whose bank pairs cover NC * 16 bytes of a row) leave unused.
      int rows[2] = {0, 0}, emax = 0, ring_row0 = 0;
      for (int l = 0; l < L; ++l) {
        const int slots = (bt.E[l] + 1) / 2 * bt.E[l];
        rows[l >> 3] = std::max(rows[l >> 3], slots);
        if (l >= 8 && l < 8 + NC) ring_row0 = std::max(ring_row0, slots);
        emax = std::max(emax, bt.E[l]);
      }
      const size_t ilv_floats = static_cast<size_t>(rows[0] + rows[1]) * 32;
      const bool ring_inside = rows[1] - ring_row0 >= 64;
      const size_t per_warp = (ilv_floats + stage_floats(scatter) + (ring_inside ? 0 : ring_floats)) * sizeof(float);   // no idle lanes
      const int ilv_warps = static_cast<int>((kScatterBudget - bt.base[L] * sizeof(uint32_t)) / per_warp);
      if (ilv_warps >= env_int("NVP_BIN_ILV_MIN_WARPS", 14) && ilv_floats <= (1u << 14) && emax < 64) {
        bt.ilv = 1; bt.ilv_rows[0] = rows[0]; bt.ilv_rows[1] = rows[1];
        region_floats = ilv_floats;
        table_bytes = bt.base[L] * sizeof(uint32_t);
        warps = std::min(cap, std::min(ilv_warps, env_int("NVP_BIN_WARPS_ILV", 32)));
        bt.dz_base = static_cast<int32_t>(ring_inside ? (rows[0] + ring_row0) * 32 : ilv_floats + stage_floats(scatter));
        bt.dz_stride = ring_inside ? 128 : NC * 16;
        bt.dz_extra_floats = ring_inside ? 0 : static_cast<int32_t>(ring_floats);
      }
    }
  }
  const int64_t avg = (n + bt.nt - 1) / bt.nt;
  int chunk = env_int("NVP_BIN_CHUNK", 0);
  if (chunk <= 0) chunk = static_cast<int>(std::min<int64_t>(1 << 20, std::max<int64_t>(128, 2 * avg)));
  bt.chunk = (chunk + 31) / 32 * 32;
  pl->bt = bt;
  pl->warps = warps;
  pl->sp_warps = sp_warps;
  pl->smem = static_cast<size_t>(warps) * (region_floats + stage_floats(scatter) + bt.dz_extra_floats) * sizeof(float) + table_bytes;
  pl->max_tasks = static_cast<int>(3 * static_cast<int64_t>(bt.nt) + 3 * n / bt.chunk + 8);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  pl->o_cnt = take(sizeof(int32_t) * 3 * bt.nt);
  pl->o_offs = take(sizeof(int32_t) * (3 * bt.nt + 1));
  pl->o_cursor = take(sizeof(int32_t) * 3 * bt.nt);
  pl->o_ntasks = take(sizeof(int32_t));
  pl->o_zeros = take(16);
  pl->o_scan = take(sizeof(unsigned long long) * (2 + (3 * static_cast<size_t>(bt.nt) + 1023) / 1024));
  pl->o_tasks = take(sizeof(int2) * pl->max_tasks);
  pl->o_recs = take(sizeof(uint4) * 3 * static_cast<size_t>(n));
  pl->total = off;
  return true;
}

void fill_bin_args(const BinPlan& pl, const LevelTab& tab, const float* coords, int64_t n, void* ws, BinArgs* a) {
  uint8_t* b = static_cast<uint8_t*>(ws);
  a->tab = tab; a->bt = pl.bt; a->coords = coords; a->n = static_cast<int32_t>(n);
  a->cnt = reinterpret_cast<int32_t*>(b + pl.o_cnt);
  a->offs = reinterpret_cast<int32_t*>(b + pl.o_offs);
  a->cursor = reinterpret_cast<int32_t*>(b + pl.o_cursor);
  a->n_tasks = reinterpret_cast<int32_t*>(b + pl.o_ntasks);
  a->scan_state = reinterpret_cast<unsigned long long*>(b + pl.o_scan);
  a->tasks = reinterpret_cast<int2*>(b + pl.o_tasks);
  a->recs = reinterpret_cast<uint4*>(b + pl.o_recs);
  a->zeros = reinterpret_cast<const uint4*>(b + pl.o_zeros);
}

int device_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

template <int F2, bool SCATTER, int THREADS, bool ILV = false>
int launch_binned_variant(const BinPlan& pl, const BinArgs& a, cudaStream_t st) {
  NVP_CUDA(cudaFuncSetAttribute(grid_binned_kernel<F2, SCATTER, THREADS, ILV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(pl.smem)));
  const int blocks = a.sp_warps > 0 ? device_sms()
                                    : std::max(1, std::min(device_sms(), (pl.max_tasks + pl.warps - 1) / pl.warps));
  grid_binned_kernel<F2, SCATTER, THREADS, ILV><<<blocks, (pl.warps + a.sp_warps) * 32, pl.smem, st>>>(a);
  return 0;
}
template <int F2, bool SCATTER>
int launch_binned_kernel(const BinPlan& pl, const BinArgs& a, cudaStream_t st) {
  // the register budget of a variant follows from its thread bound (64 K registers per SM)
  const int warps = pl.warps + a.sp_warps;
  if constexpr (SCATTER && F2 == 2) {
    if (a.bt.ilv) {
      if (warps <= 16) return launch_binned_variant<F2, SCATTER, 512, true>(pl, a, st);
      if (warps <= 18) return launch_binned_variant<F2, SCATTER, 576, true>(pl, a, st);
      return launch_binned_variant<F2, SCATTER, 768, true>(pl, a, st);
    }
  }
  if (warps <= 16) return launch_binned_variant<F2, SCATTER, 512>(pl, a, st);
  if (warps <= 24) return launch_binned_variant<F2, SCATTER, 768>(pl, a, st);
  return launch_binned_variant<F2, SCATTER, 1024>(pl, a, st);
}
template <bool SCATTER>
int launch_binned(int f2, const BinPlan& pl, const BinArgs& a, cudaStream_t st) {
  switch (f2) {
    case 1: return launch_binned_kernel<1, SCATTER>(pl, a, st);
    case 2: return launch_binned_kernel<2, SCATTER>(pl, a, st);
    case 4: return launch_binned_kernel<4, SCATTER>(pl, a, st);
    case 8: return launch_binned_kernel<8, SCATTER>(pl, a, st);
  }
  set_error("n_features_per_level must be 1, 2, 4 or 8");
  return 1;
}

}  // namespace

int launch_grid_gather(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords, int64_t n,
                       float* z, int ldz, uint8_t* z16t, int kz, cudaStream_t st, bool temporal_interp) {
  GridArgs a{};
  a.tab = tab;
  a.coords = coords;
  a.n = n;
  a.kf[0] = p->kf_xy; a.kf[1] = p->kf_yt; a.kf[2] = p->kf_xt;
  a.sparse = p->sparse;
  a.z = z; a.ldz = ldz; a.z16t = z16t; a.kz = kz; a.n_pad = (n + 127) / 128 * 128;
  a.tres = d->t_resolution; a.xres = d->x_resolution; a.yres = d->y_resolution;
  a.interp = temporal_interp ? 1 : 0;
  a.scale = 1.0f;
  if (z == nullptr && z16t != nullptr && (tab.n_levels * d->n_features) % 8 == 0 && n >= 4096) {
    switch (d->n_features) {
      case 1: return launch_gather_v2<1>(a, d->sparse_features, st);
      case 2: return launch_gather_v2<2>(a, d->sparse_features, st);
      case 4: return launch_gather_v2<4>(a, d->sparse_features, st);
      case 8: return launch_gather_v2<8>(a, d->sparse_features, st);
    }
  }
  return dispatch(false, d->n_features, d->sparse_features, a, st);
}

int launch_grid_scatter(const nvp_desc* d, const LevelTab& tab, const float* coords, int64_t n, const float* dz,
                        int lddz, const uint8_t* dz16t, int kz, float scale, const float* scale_ptr, const nvp_grads* g,
                        cudaStream_t st) {
  if (!g->kf_xy && !g->kf_yt && !g->kf_xt && !g->sparse) return 0;
  GridArgs a{};
  a.tab = tab;
  a.coords = coords;
  a.n = n;
  a.gkf[0] = g->kf_xy; a.gkf[1] = g->kf_yt; a.gkf[2] = g->kf_xt;
  a.gsparse = g->sparse;
  a.z = const_cast<float*>(dz); a.ldz = lddz;
  a.dz16t = dz16t; a.kz = kz;
  a.tres = d->t_resolution; a.xres = d->x_resolution; a.yres = d->y_resolution;
  a.scale = scale;
  a.scale_ptr = scale_ptr;
  // Coarse levels -> shared-memory privatised kernel (as many levels as fit ~150 KB for three planes).
  a.lvl_begin = 0;
  const bool any_kf = g->kf_xy || g->kf_yt || g->kf_xt;
  if (any_kf && n >= 4096) {
    int lc = 0;
    while (lc < tab.n_levels && static_cast<size_t>(tab.offset[lc + 1]) * d->n_features * 3 * sizeof(float) <= 150 * 1024) ++lc;
    if (lc > 0) {
      a.n_coarse = lc;
      a.coarse_cells = tab.offset[lc];
      a.lvl_begin = lc;
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const int blocks = static_cast<int>(std::min<int64_t>(sms, (n + 1023) / 1024));
      a.chunk = (n + blocks - 1) / blocks;
      const size_t smem = static_cast<size_t>(a.coarse_cells) * d->n_features * 3 * sizeof(float);
      ScopedKernelTimer timer(K_SCATTER, st);
      int rc = 1;
      switch (d->n_features) {
        case 1: rc = launch_coarse<1>(a, blocks, smem, st); break;
        case 2: rc = launch_coarse<2>(a, blocks, smem, st); break;
        case 4: rc = launch_coarse<4>(a, blocks, smem, st); break;
        case 8: rc = launch_coarse<8>(a, blocks, smem, st); break;
      }
      if (rc) return rc;
      NVP_LAUNCH_CHECK();
    }
  }
  return dispatch(true, d->n_features, d->sparse_features, a, st);
}

// ---- binned path (see grid_binned.cuh) ---------------------------------------------------------
size_t grid_bin_workspace_bytes(const nvp_desc* d, const LevelTab& tab, int64_t n) {
  BinPlan pl;
  return plan_bins(d, tab, n, &pl) ? pl.total + 256 : 0;
}

void grid_bin_plan_info(const nvp_desc* d, const LevelTab& tab, int64_t n, int* tb, int* chunk, int32_t* extent, int32_t* base,
                        size_t* workspace) {
  BinPlan pl;
  const bool ok = plan_bins(d, tab, n, &pl);
  if (tb) *tb = ok ? pl.bt.tb : 0;
  if (chunk) *chunk = ok ? pl.bt.chunk : 0;
  if (workspace) *workspace = ok ? pl.total + 256 : 0;
  for (int l = 0; ok && l < tab.n_levels; ++l) {
    if (extent) extent[l] = pl.bt.E[l];
    if (base) base[l] = pl.bt.base[l];
  }
  if (ok && base) base[tab.n_levels] = pl.bt.base[tab.n_levels];
}

int launch_grid_bin(const nvp_desc* d, const LevelTab& tab, const float* coords, int64_t n, int kz, void* binws,
                    cudaStream_t st) {
  BinPlan pl;
  NVP_CHECK(binws != nullptr && plan_bins(d, tab, n, &pl), "binned grid path not available for this configuration");
  BinArgs a{};
  fill_bin_args(pl, tab, coords, n, binws, &a);
  a.kz = kz;
  // counters, and (contiguous in the workspace) offs / cursor / n_tasks / the 16 zero bytes / the scan's chain state
  NVP_CUDA(cudaMemsetAsync(a.cnt, 0, pl.o_tasks, st));
  const int blocks = static_cast<int>(std::min<int64_t>(8 * device_sms(), (n + 255) / 256));
  ScopedKernelTimer timer(K_BIN, st);
  grid_bin_count_kernel<<<blocks, 256, 0, st>>>(a);
  NVP_LAUNCH_CHECK();
  grid_bin_scan_kernel<<<(3 * pl.bt.nt + 1023) / 1024, 1024, 0, st>>>(a);
  NVP_LAUNCH_CHECK();
  grid_bin_fill_kernel<<<blocks, 256, 0, st>>>(a);
  NVP_LAUNCH_CHECK();
  return 0;
}

int launch_grid_gather_binned(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords,
                              int64_t n, uint8_t* z16t, int kz, void* binws, cudaStream_t st, bool temporal_interp) {
  BinPlan pl;
  NVP_CHECK(binws != nullptr && plan_bins(d, tab, n, &pl), "binned grid path not available for this configuration");
  BinArgs b{};
  fill_bin_args(pl, tab, coords, n, binws, &b);
  b.kf[0] = p->kf_xy; b.kf[1] = p->kf_yt; b.kf[2] = p->kf_xt;
  b.z16t = z16t; b.kz = kz;
  // 3x3 voxel neighbourhood + padding columns (+ all-zero padding rows)
  GridArgs a{};
  a.tab = tab; a.coords = coords; a.n = n;
  a.sparse = p->sparse;
  a.z16t = z16t; a.kz = kz; a.n_pad = (n + 127) / 128 * 128;
  a.tres = d->t_resolution; a.xres = d->x_resolution; a.yres = d->y_resolution;
  a.interp = temporal_interp ? 1 : 0;
  a.scale = 1.0f;
  a.n_coarse_chunks = tab.n_levels * d->n_features / 8;   // no keyframe passes in gather_fine_kernel
  a.zero_planes = 1;
  b.sp = a; b.sp_col0 = 3 * tab.n_levels * d->n_features;
  b.win_warps = pl.warps; b.sp_warps = pl.sp_warps;
  {
    ScopedKernelTimer timer(K_GATHER, st);
    if (int rc = launch_binned<false>(d->n_features, pl, b, st)) return rc;
    NVP_LAUNCH_CHECK();
  }
  if (pl.sp_warps > 0) return 0;
  dim3 grid(static_cast<unsigned>((a.n_pad + kGridThreads - 1) / kGridThreads), 1);
  ScopedKernelTimer timer(K_GATHER, st);
  switch (d->n_features * 16 + d->sparse_features) {
#define NVP_CASE(F2, F3) case F2 * 16 + F3: gather_fine_kernel<F2, F3><<<grid, kGridThreads, 0, st>>>(a); break;
    NVP_CASE(1, 1) NVP_CASE(1, 2) NVP_CASE(1, 4) NVP_CASE(1, 8) NVP_CASE(2, 1) NVP_CASE(2, 2) NVP_CASE(2, 4) NVP_CASE(2, 8)
    NVP_CASE(4, 1) NVP_CASE(4, 2) NVP_CASE(4, 4) NVP_CASE(4, 8) NVP_CASE(8, 1) NVP_CASE(8, 2) NVP_CASE(8, 4) NVP_CASE(8, 8)
#undef NVP_CASE
    default: NVP_CHECK(false, "n_features_per_level must be 1, 2, 4 or 8");
  }
  NVP_LAUNCH_CHECK();
  return 0;
}

int launch_grid_scatter_binned(const nvp_desc* d, const LevelTab& tab, const float* coords, int64_t n,
                               const uint8_t* dz16t, int kz, float scale, const float* scale_ptr, const nvp_grads* g,
                               void* binws, cudaStream_t st) {
  if (!g->kf_xy && !g->kf_yt && !g->kf_xt && !g->sparse) return 0;
  BinPlan pl;
  NVP_CHECK(binws != nullptr && plan_bins(d, tab, n, &pl, true), "binned grid path not available for this configuration");
  GridArgs a{};
  a.tab = tab; a.coords = coords; a.n = n;
  a.gsparse = g->sparse;
  a.dz16t = dz16t; a.kz = kz;
  a.tres = d->t_resolution; a.xres = d->x_resolution; a.yres = d->y_resolution;
  a.scale = scale; a.scale_ptr = scale_ptr;
  a.lvl_begin = tab.n_levels;
  const bool any_kf = g->kf_xy || g->kf_yt || g->kf_xt;
  if (any_kf) {
    BinArgs b{};
    fill_bin_args(pl, tab, coords, n, binws, &b);
    b.gkf[0] = g->kf_xy; b.gkf[1] = g->kf_yt; b.gkf[2] = g->kf_xt;
    b.z16t = const_cast<uint8_t*>(dz16t); b.kz = kz;
    b.scale = scale; b.scale_ptr = scale_ptr;
    b.sp = a; b.sp_col0 = 3 * tab.n_levels * d->n_features;
    b.win_warps = pl.warps; b.sp_warps = pl.sp_warps;
    ScopedKernelTimer timer(K_SCATTER, st);
    if (int rc = launch_binned<true>(d->n_features, pl, b, st)) return rc;
    NVP_LAUNCH_CHECK();
    if (pl.sp_warps > 0) return 0;
  }
  if (g->sparse == nullptr) return 0;
  const int col0 = 3 * tab.n_levels * d->n_features;
  const unsigned blocks = static_cast<unsigned>((3 * n + 255) / 256);
  ScopedKernelTimer timer(K_SCATTER, st);
  switch (d->sparse_features) {
    case 1: sparse_scatter_rows_kernel<1><<<blocks, 256, 0, st>>>(a, col0); break;
    case 2: sparse_scatter_rows_kernel<2><<<blocks, 256, 0, st>>>(a, col0); break;
    case 4: sparse_scatter_rows_kernel<4><<<blocks, 256, 0, st>>>(a, col0); break;
    case 8: sparse_scatter_rows_kernel<8><<<blocks, 256, 0, st>>>(a, col0); break;
    default: NVP_CHECK(false, "3d n_features_per_level must be 1, 2, 4 or 8");
  }
  NVP_LAUNCH_CHECK();
  return 0;
}


}  // namespace nvp
