// Tile-binned keyframe gather / scatter-add (included by grid.cu; tensor-core path only).
//
// The direct kernels in grid.cu issue one scattered 8-byte access per (sample, level, corner): 192 per sample and
// direction.  Forward that is bound by the LSU's sector rate; backward by the per-SM rate of global reductions
// (measured 1.29 cycles per lane-level red.global on this part: 240 M of them = 1.1 ms of a 4.7 ms step).
//
// Here the samples of a call are first bucketed, per keyframe plane, by the TB x TB tile of the unit square their
// plane coordinates (u0, u1) fall into (grid_bin_* kernels: histogram, scan + task list, fill; TB = 128 by default).
// All samples of one tile touch, at level l, only the cells of a small window: E_l = ceil(scale_l / TB) + 2 cells per
// axis (619 cells = 5 KB for all 16 levels of config S).  One WARP owns one task (a tile, or a chunk of a crowded tile)
// and a private window region in shared memory:
//   gather : the window is copied from the plane table with cp.async (rows coalesced, all cells in flight together);
//            then lane = sample: every lane walks the levels, reads its 4 corners from the window and writes its slice
//            of the latent row as whole 16-byte chunks - same corner order and fma chain as the direct kernel;
//   scatter: lane = (level, cell row): the region starts at zero, the warp takes one sample per step and every lane does
//            a plain (non-atomic) read-modify-write of two adjacent corners - lanes of one sample never collide
//            (different levels / rows), consecutive samples are ordered by __syncwarp - and the region is flushed once
//            per task with one vector reduction per non-zero cell: ~8 global reductions per (sample, plane), not 64.
// Cell addressing keeps the reference's edge behaviour (flat index without clamping, then modulo the level size,
// SURVEY.md A.2): the window stores the "virtual" cells (res, j) / (i, res) and maps them to their aliases on load /
// flush.  Samples whose cells fall outside the window (coordinates outside [0,1]) take the direct global path.
// The 3x3 voxel neighbourhood of the 3-D grid (DRAM-latency-bound) runs as a few extra warps of the same CTAs.
// tests/test_binned_algorithm.py restates this algorithm in numpy and checks it against the oracle on CPU.
#pragma once
// (no namespace of its own: grid.cu includes this file inside nvp::<anonymous>, after the helpers it uses)

struct BinTab {
  int32_t E[NVP_MAX_LEVELS];         // window extent per axis, in cells
  int32_t base[NVP_MAX_LEVELS + 1];  // first region cell of level l; base[L] = cells per region
  uint32_t magic[NVP_MAX_LEVELS];    // idx / E == (idx * magic) >> 20 for idx < E * E
  int32_t tb, log_tb, nt;            // tiles per axis (power of two), log2, tb * tb
  int32_t chunk;                     // samples per task
};

struct BinArgs {
  LevelTab tab;
  BinTab bt;
  const float* coords;
  int32_t n;
  // bucket state (caller workspace)
  int32_t* cnt;            // [3 * nt]     samples per (plane, tile)
  int32_t* offs;           // [3 * nt + 1] exclusive prefix of cnt (positions into recs)
  int32_t* cursor;         // [3 * nt]     fill cursors
  int2* tasks;             // [max_tasks]  (bucket, first position)
  int32_t* n_tasks;        // [1]
  uint4* recs;             // [3 * n]      StagedSample records grouped by bucket (written by the fill kernel)
  // tables
  const float* kf[3];
  float* gkf[3];
  uint8_t* z16t;           // gather out / scatter in (fp16 MMA tile format)
  int kz;
  float scale;
  const float* scale_ptr;
  const uint4* zeros;      // 16 zero bytes (source of padding entries' latent gradient)
  // voxel-neighbourhood role: warps [win_warps, blockDim/32) of every CTA run the 3x3-voxel gather (+ padding
  // columns) / scatter-add of grid.cu next to the window warps - DRAM-latency-bound work hidden under the
  // shared-memory-bound window work.  sp_warps == 0: separate kernels do it.
  GridArgs sp;
  int sp_col0;             // first voxel column of the latent row (3 * L * F2)
  int win_warps, sp_warps;
};

constexpr int kBinThreadsMax = 1024;   // kernel variants are compiled for 512 / 768 / 1024 threads per CTA

// One bucket entry: byte offset of the sample's latent row inside the tile buffer (0xffffffff = padding entry of a
// partial batch), the row's swizzle term ((row & 7) << 4) and the sample's two plane coordinates.
struct __align__(16) StagedSample { uint32_t rowoff, swz; float u0, u1; };
// Per-warp scratch behind the window region: the scatter-add's staged batch (32 x 16 B) or the gather's per-task level
// table (n_levels <= 32 x 32 B).
struct __align__(16) LevelWindow { float scale; int lo0, lo1, base; int E, amax0, amax1, res; };
constexpr int kStageFloats = NVP_MAX_LEVELS * sizeof(LevelWindow) / sizeof(float);
static_assert(kStageFloats * sizeof(float) >= 32 * sizeof(StagedSample), "stage too small");

__device__ __forceinline__ int bin_tile_axis(float u, int tb) {
  const int b = __float2int_rz(u * static_cast<float>(tb));   // tb is a power of two: the product is exact
  return min(max(b, 0), tb - 1);
}
// plane order xy | yt | xt with inputs (x, y), (t, y), (t, x)   (modules.py:61-63)
__device__ __forceinline__ int bin_bucket(const BinTab& bt, int plane, float t, float x, float y) {
  const float u0 = plane == 0 ? x : t, u1 = plane == 2 ? x : y;
  return plane * bt.nt + (bin_tile_axis(u1, bt.tb) << bt.log_tb) + bin_tile_axis(u0, bt.tb);
}

__global__ void __launch_bounds__(256) grid_bin_count_kernel(const BinArgs a) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < a.n; s += gridDim.x * blockDim.x) {
    const float t = __ldg(a.coords + 3 * static_cast<int64_t>(s)), x = __ldg(a.coords + 3 * static_cast<int64_t>(s) + 1),
                y = __ldg(a.coords + 3 * static_cast<int64_t>(s) + 2);
#pragma unroll
    for (int p = 0; p < 3; ++p) atomicAdd(a.cnt + bin_bucket(a.bt, p, t, x, y), 1);
  }
}

// One CTA: exclusive prefix sums of the bucket sizes (-> offs, cursor) and of the per-bucket task counts, then the
// task list itself.  A bucket of c samples becomes ceil(c / chunk) tasks.  The counters are staged in shared memory
// with coalesced reads; each thread then owns a contiguous run of buckets.
__global__ void __launch_bounds__(1024) grid_bin_scan_kernel(const BinArgs a) {
  extern __shared__ __align__(16) int s_cnt[];   // [3 * nt]
  __shared__ int s_c[1024], s_t[1024];
  const int m = 3 * a.bt.nt;
  {
    // coalesced 16-byte loads, several in flight per thread (m = 3 * tb^2 is a multiple of 4)
    const int4* src = reinterpret_cast<const int4*>(a.cnt);
    int4* dst = reinterpret_cast<int4*>(s_cnt);
#pragma unroll 4
    for (int i = threadIdx.x; i < m / 4; i += 1024) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const int per = ((m + 1023) / 1024) | 1;   // odd run length: the threads' strided walks hit distinct banks
  const int i0 = min(m, static_cast<int>(threadIdx.x) * per), i1 = min(m, i0 + per);
  int csum = 0, tsum = 0;
  for (int i = i0; i < i1; ++i) {
    const int c = s_cnt[i];
    csum += c;
    tsum += c == 0 ? 0 : (c <= a.bt.chunk ? 1 : (c + a.bt.chunk - 1) / a.bt.chunk);
  }
  s_c[threadIdx.x] = csum; s_t[threadIdx.x] = tsum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {   // Hillis-Steele inclusive scan
    const int vc = threadIdx.x >= d ? s_c[threadIdx.x - d] : 0, vt = threadIdx.x >= d ? s_t[threadIdx.x - d] : 0;
    __syncthreads();
    s_c[threadIdx.x] += vc; s_t[threadIdx.x] += vt;
    __syncthreads();
  }
  int co = s_c[threadIdx.x] - csum, to = s_t[threadIdx.x] - tsum;
  for (int i = i0; i < i1; ++i) {
    const int c = s_cnt[i];
    s_cnt[i] = co;
    for (int b = 0; b < c; b += a.bt.chunk) a.tasks[to++] = make_int2(i, co + b);
    co += c;
  }
  __syncthreads();
  {
    int4* o4 = reinterpret_cast<int4*>(a.offs);
    int4* c4 = reinterpret_cast<int4*>(a.cursor);
    const int4* s4 = reinterpret_cast<const int4*>(s_cnt);
#pragma unroll 4
    for (int i = threadIdx.x; i < m / 4; i += 1024) { const int4 o = s4[i]; o4[i] = o; c4[i] = o; }
  }
  if (threadIdx.x == 1023) { a.offs[m] = s_c[1023]; a.n_tasks[0] = s_t[1023]; }
}

__global__ void __launch_bounds__(256) grid_bin_fill_kernel(const BinArgs a) {
  const uint32_t tile_bytes = static_cast<uint32_t>(a.kz) * tc::kPanelBytes;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < a.n; s += gridDim.x * blockDim.x) {
    const float t = __ldg(a.coords + 3 * static_cast<int64_t>(s)), x = __ldg(a.coords + 3 * static_cast<int64_t>(s) + 1),
                y = __ldg(a.coords + 3 * static_cast<int64_t>(s) + 2);
    const uint32_t r = static_cast<uint32_t>(s) & 127u;
    const uint32_t rowoff = static_cast<uint32_t>(s >> 7) * tile_bytes + r * 128u, swz = (r & 7u) << 4;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const float u0 = p == 0 ? x : t, u1 = p == 2 ? x : y;
      a.recs[atomicAdd(a.cursor + bin_bucket(a.bt, p, t, x, y), 1)] =
          make_uint4(rowoff, swz, __float_as_uint(u0), __float_as_uint(u1));
    }
  }
}

// ---- shared-memory cell access ----------------------------------------------------------------
template <int F>
__device__ __forceinline__ void lds_feat(const float* p, float (&v)[F]) {
  if constexpr (F == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  } else if constexpr (F == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (F == 8) {
    const float4 t = *reinterpret_cast<const float4*>(p), u = *(reinterpret_cast<const float4*>(p) + 1);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; v[4] = u.x; v[5] = u.y; v[6] = u.z; v[7] = u.w;
  } else {
#pragma unroll
    for (int f = 0; f < F; ++f) v[f] = p[f];
  }
}
template <int F>
__device__ __forceinline__ void sts_feat(float* p, const float (&v)[F]) {
  if constexpr (F == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else if constexpr (F == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (F == 8) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *(reinterpret_cast<float4*>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int f = 0; f < F; ++f) p[f] = v[f];
  }
}

// Rare path (plane coordinates outside [0,1], i.e. cells outside the task's window): the direct global access of the
// unbinned kernels.  Out of line to keep the hot loops small.
template <int F2> struct CornerPair { float a[F2], b[F2]; };
template <int F2>
__device__ __noinline__ CornerPair<F2> direct_corner_pair_load(const float* __restrict__ kfl, int flat, int cells) {
  CornerPair<F2> r;
  ld_feat<F2>(kfl + static_cast<size_t>(wrap_cell(flat, cells)) * F2, r.a);
  ld_feat<F2>(kfl + static_cast<size_t>(wrap_cell(flat + 1, cells)) * F2, r.b);
  return r;
}
template <int F2>
__device__ __noinline__ void direct_corner_pair_add(float* __restrict__ gkl, int flat, int cells, float ka, float kb,
                                                    CornerPair<F2> d) {   // d.a = the sample's latent gradient
  float a[F2], b[F2];
#pragma unroll
  for (int f = 0; f < F2; ++f) { a[f] = ka * d.a[f]; b[f] = kb * d.a[f]; }
  red_feat<F2>(gkl + static_cast<size_t>(wrap_cell(flat, cells)) * F2, a);
  red_feat<F2>(gkl + static_cast<size_t>(wrap_cell(flat + 1, cells)) * F2, b);
}

// Asynchronous global -> shared copy of one cell (cp.async, LDGSTS in SASS): no register round trip, so all cells of
// a window are in flight together.  src_bytes = 0 zero-fills (cells outside the table).
template <int BYTES>
__device__ __forceinline__ void cp_async_cell(float* smem_dst, const float* gmem_src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(tc::smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Window <-> table transfer for levels [lb, le): LOAD copies the window into the region (asynchronously: the caller
// waits), FLUSH adds the region (times `scale`) into the gradient table.
enum { REGION_LOAD = 0, REGION_FLUSH = 2 };

template <int F2, int OP>
__device__ __forceinline__ void region_io(float* __restrict__ reg, const float* __restrict__ src, float* __restrict__ dst,
                                          const float* s_scale, const int* s_res, const int* s_off, const int* s_E,
                                          const int* s_base, const uint32_t* s_magic, int lb, int le, float ub0,
                                          float ub1, float scale, int lane) {
  for (int l = lb; l < le; ++l) {
    const int E = s_E[l], cnt = E * E, res = s_res[l];
    float* rl = reg + static_cast<size_t>(s_base[l]) * F2;
    const float sc = s_scale[l];
    const int lo0 = static_cast<int>(floorf(fmaf(sc, ub0, 0.5f))), lo1 = static_cast<int>(floorf(fmaf(sc, ub1, 0.5f)));
    const int cells = res * res;
    const uint32_t magic = s_magic[l];
    const size_t goff = static_cast<size_t>(s_off[l]);
    for (int idx = lane; idx < cnt; idx += 32) {
      const int bb = static_cast<int>((static_cast<uint32_t>(idx) * magic) >> 20), aa = idx - bb * E;
      const int g0 = lo0 + aa, g1 = lo1 + bb;
      const bool ok = g0 <= res && g1 <= res;   // (res, j) and (i, res) are the aliased "virtual" cells
      const size_t cell = ok ? goff + wrap_cell(g0 + g1 * res, cells) : goff;
      if constexpr (OP == REGION_LOAD) {
        constexpr int kBytes = F2 * 4 < 16 ? F2 * 4 : 16;
#pragma unroll
        for (int b = 0; b < F2 * 4; b += kBytes)
          cp_async_cell<kBytes>(rl + idx * F2 + b / 4, src + cell * F2 + b / 4, ok ? kBytes : 0);
      } else {
        float v[F2];
        lds_feat<F2>(rl + idx * F2, v);
        bool nz = false;
#pragma unroll
        for (int f = 0; f < F2; ++f) { nz |= v[f] != 0.0f; v[f] *= scale; }
        if (ok && nz) red_feat<F2>(dst + cell * F2, v);
      }
    }
  }
}

template <int F2, bool SCATTER, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) grid_binned_kernel(const BinArgs a) {
  extern __shared__ __align__(16) float s_region[];
  __shared__ float s_scale[NVP_MAX_LEVELS];
  __shared__ int s_res[NVP_MAX_LEVELS], s_off[NVP_MAX_LEVELS], s_E[NVP_MAX_LEVELS], s_base[NVP_MAX_LEVELS + 1];
  __shared__ uint32_t s_magic[NVP_MAX_LEVELS];
  const int L = a.tab.n_levels;
  if (threadIdx.x < L) {
    s_scale[threadIdx.x] = a.tab.scale[threadIdx.x];
    s_res[threadIdx.x] = a.tab.res[threadIdx.x];
    s_off[threadIdx.x] = a.tab.offset[threadIdx.x];
    s_E[threadIdx.x] = a.bt.E[threadIdx.x];
    s_magic[threadIdx.x] = a.bt.magic[threadIdx.x];
  }
  if (threadIdx.x <= L) s_base[threadIdx.x] = a.bt.base[threadIdx.x];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = a.win_warps;
  if (warp >= nwarps) {
    if constexpr (F2 <= 4) {   // (the role is only enabled when the 3-D grid has F2 features per voxel)
      const int64_t gtid = (static_cast<int64_t>(blockIdx.x) * a.sp_warps + (warp - nwarps)) * 32 + lane;
      const int64_t gstride = static_cast<int64_t>(gridDim.x) * a.sp_warps * 32;
      if constexpr (SCATTER) {
        if (a.sp.gsparse != nullptr) {
          const float sscale = a.sp.scale_ptr ? a.sp.scale * __ldg(a.sp.scale_ptr) : a.sp.scale;
#pragma unroll 2
          for (int64_t idx = gtid; idx < 3 * a.sp.n; idx += gstride) sparse_scatter_row<F2>(a.sp, a.sp_col0, sscale, idx);
        }
      } else {
#pragma unroll 2
        for (int64_t s = gtid; s < a.sp.n_pad; s += gstride) sparse_pad_sample<F2>(a.sp, a.sp_col0, s);
      }
    }
    return;
  }
  const int region_floats = (a.bt.base[L] * F2 + 3) & ~3;
  float* reg = s_region + static_cast<size_t>(warp) * (region_floats + kStageFloats);
  StagedSample* stage = reinterpret_cast<StagedSample*>(reg + region_floats);
  const int n_tasks = __ldg(a.n_tasks);
  const float inv_tb = 1.0f / static_cast<float>(a.bt.tb);
  const int pw = L * F2;
  float scale = 1.0f;
  if constexpr (SCATTER) scale = a.scale_ptr ? a.scale * __ldg(a.scale_ptr) : a.scale;
  const int c1 = lane & 1;

  for (int task = blockIdx.x * nwarps + warp; task < n_tasks; task += gridDim.x * nwarps) {
    const int2 tk = __ldg(a.tasks + task);
    const int bucket = tk.x, beg = tk.y;
    const int end = min(beg + a.bt.chunk, __ldg(a.offs + bucket + 1));
    const int plane = bucket >> (2 * a.bt.log_tb), tile = bucket & (a.bt.nt - 1);
    const float ub0 = static_cast<float>(tile & (a.bt.tb - 1)) * inv_tb, ub1 = static_cast<float>(tile >> a.bt.log_tb) * inv_tb;
    const float* __restrict__ kfp = a.kf[plane];
    float* __restrict__ gkp = a.gkf[plane];
    const uint4* __restrict__ recs = a.recs;
    if (SCATTER && gkp == nullptr) continue;
    const uint8_t* zero_src = reinterpret_cast<const uint8_t*>(a.zeros);
    const uint4 pad_rec = make_uint4(0xffffffffu, 0u, __float_as_uint(ub0), __float_as_uint(ub1));

    if constexpr (!SCATTER) {
      // ---- gather: lane = sample.  Every lane walks the levels itself (4 corner reads per level from the warp's
      // window), so the per-sample overhead (record, addressing, output row) is paid once per 32 samples and a
      // latent row leaves as whole 16-byte chunks.  Same corner order / fma chain as the direct kernel.
      LevelWindow* lw = reinterpret_cast<LevelWindow*>(stage);
      if (lane < L) {
        LevelWindow q;
        q.scale = s_scale[lane]; q.res = s_res[lane]; q.E = s_E[lane];
        q.lo0 = static_cast<int>(floorf(fmaf(q.scale, ub0, 0.5f)));
        q.lo1 = static_cast<int>(floorf(fmaf(q.scale, ub1, 0.5f)));
        q.base = s_base[lane] * F2;
        q.amax0 = min(q.E - 2, q.res - 1 - q.lo0); q.amax1 = min(q.E - 2, q.res - 1 - q.lo1);
        lw[lane] = q;
      }
      region_io<F2, REGION_LOAD>(reg, kfp, nullptr, s_scale, s_res, s_off, s_E, s_base, s_magic, 0, L, ub0, ub1, 1.0f, lane);
      cp_async_wait_all();
      __syncwarp();
      constexpr int LPC = 8 / F2;                       // levels per 16-byte chunk of the latent row
      const int col0 = plane * pw;
      uint4 rec = beg + lane < end ? __ldg(recs + beg + lane) : pad_rec;
      for (int i = beg; i < end; i += 32) {
        const uint4 cur = rec;
        rec = i + 32 + lane < end ? __ldg(recs + i + 32 + lane) : pad_rec;
        const float u0 = __uint_as_float(cur.z), u1 = __uint_as_float(cur.w);
        const bool valid = cur.x != 0xffffffffu;
        for (int ch = 0; ch < L / LPC; ++ch) {
          float o[8];
#pragma unroll
          for (int j = 0; j < LPC; ++j) {
            const int l = ch * LPC + j;
            const LevelWindow q = lw[l];
            const float p0 = fmaf(q.scale, u0, 0.5f), p1 = fmaf(q.scale, u1, 0.5f);
            const float f0 = floorf(p0), f1 = floorf(p1);
            const int i0 = static_cast<int>(f0), i1 = static_cast<int>(f1);
            const float w0 = p0 - f0, w1 = p1 - f1;
            const int aa = i0 - q.lo0, bb = i1 - q.lo1;
            float v00[F2], v10[F2], v01[F2], v11[F2];
            if (static_cast<unsigned>(aa) <= static_cast<unsigned>(q.amax0) && static_cast<unsigned>(bb) <= static_cast<unsigned>(q.amax1)) {
              const float* cp = reg + q.base + (bb * q.E + aa) * F2;
              lds_feat<F2>(cp, v00);
              lds_feat<F2>(cp + F2, v10);
              lds_feat<F2>(cp + q.E * F2, v01);
              lds_feat<F2>(cp + q.E * F2 + F2, v11);
            } else {
              const float* kfl = kfp + static_cast<size_t>(s_off[l]) * F2;
              const CornerPair<F2> r0 = direct_corner_pair_load<F2>(kfl, i0 + i1 * q.res, q.res * q.res);
              const CornerPair<F2> r1 = direct_corner_pair_load<F2>(kfl, i0 + (i1 + 1) * q.res, q.res * q.res);
#pragma unroll
              for (int f = 0; f < F2; ++f) { v00[f] = r0.a[f]; v10[f] = r0.b[f]; v01[f] = r1.a[f]; v11[f] = r1.b[f]; }
            }
            const float a0 = 1.0f - w0, a1 = 1.0f - w1;
            const float k00 = a0 * a1, k10 = w0 * a1, k01 = a0 * w1, k11 = w0 * w1;
#pragma unroll
            for (int f = 0; f < F2; ++f) {
              float r = k00 * v00[f];
              r = fmaf(k10, v10[f], r);
              r = fmaf(k01, v01[f], r);
              r = fmaf(k11, v11[f], r);
              o[j * F2 + f] = r;
            }
          }
          if (valid) {
            const int col = col0 + ch * 8;
            uint8_t* dst = a.z16t + static_cast<size_t>(col >> 6) * tc::kPanelBytes + cur.x +
                           ((static_cast<uint32_t>((col & 63) >> 3) << 4) ^ cur.y);
            *reinterpret_cast<uint4*>(dst) = make_uint4(tc::pack_half2(o[0], o[1]), tc::pack_half2(o[2], o[3]),
                                                        tc::pack_half2(o[4], o[5]), tc::pack_half2(o[6], o[7]));
          }
        }
      }
      __syncwarp();   // the level table and the window are rewritten by the next task
    } else for (int lb = 0; lb < L; lb += 16) {
      // ---- scatter-add: lane = (level, cell row); one sample per step of the warp
      const int le = min(L, lb + 16);
      const bool lv = lb + (lane >> 1) < L;
      const int l = lv ? lb + (lane >> 1) : L - 1;
      // per-lane constants of this task: its level's window
      const float sc = s_scale[l];
      const int res = s_res[l], cells = res * res, E = s_E[l];
      const size_t goff = static_cast<size_t>(s_off[l]);
      const int lo0 = static_cast<int>(floorf(fmaf(sc, ub0, 0.5f))), lo1 = static_cast<int>(floorf(fmaf(sc, ub1, 0.5f)));
      const unsigned amax0 = static_cast<unsigned>(min(E - 2, res - 1 - lo0)), amax1 = static_cast<unsigned>(min(E - 2, res - 1 - lo1));
      float* rl = reg + (static_cast<size_t>(s_base[l]) + c1 * E) * F2;
      const int col = plane * pw + l * F2;
      uint8_t* zb = a.z16t + static_cast<size_t>(col >> 6) * tc::kPanelBytes + ((col & 7) << 1);
      const uint32_t cc16 = static_cast<uint32_t>((col & 63) >> 3) << 4;

      uint4 rec = beg + lane < end ? __ldg(recs + beg + lane) : pad_rec;
      {
        float4* r4 = reinterpret_cast<float4*>(reg);
        for (int i = lane; i < region_floats / 4; i += 32) r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();

      // one sample: window-relative cell of the lane's level, the two corners of the lane's row and their weights
      auto geometry = [&](const StagedSample& q, float& ka, float& kb, int& i0, int& i1, int& loc) -> bool {
        const float p0 = fmaf(sc, q.u0, 0.5f), p1 = fmaf(sc, q.u1, 0.5f);
        const float f0 = floorf(p0), f1 = floorf(p1);
        i0 = static_cast<int>(f0); i1 = static_cast<int>(f1);
        const float w0 = p0 - f0, w1 = p1 - f1;
        const float wr = c1 ? w1 : 1.0f - w1;
        ka = (1.0f - w0) * wr; kb = w0 * wr;
        const int aa = i0 - lo0, bb = i1 - lo1;
        loc = (bb * E + aa) * F2;
        return static_cast<unsigned>(aa) <= amax0 && static_cast<unsigned>(bb) <= amax1;
      };

      for (int i = beg; i < end; i += 32) {
        reinterpret_cast<uint4*>(stage)[lane] = rec;
        __syncwarp();
        rec = i + 32 + lane < end ? __ldg(recs + i + 32 + lane) : pad_rec;   // lands while this batch is processed
        const int cnt = min(32, end - i);

        if constexpr (SCATTER) {
          // latent-gradient values are fetched (as raw bits) one group ahead of their use: two register buffers
          constexpr int G = F2 <= 2 ? 8 : (F2 == 4 ? 4 : 2);
          constexpr int NG = 32 / G;
          RawHalfs<F2> dA[G], dB[G];
          auto load_group = [&](int g, RawHalfs<F2> (&dg)[G]) {
#pragma unroll
            for (int j = 0; j < G; ++j) {
              const StagedSample q = stage[g * G + j];
              const uint8_t* src = q.rowoff != 0xffffffffu ? zb + q.rowoff + (cc16 ^ q.swz) : zero_src;
              dg[j] = ld_halfs_raw<F2>(src);
            }
          };
          auto process_group = [&](int g, const RawHalfs<F2> (&dg)[G]) {
#pragma unroll
            for (int j = 0; j < G; ++j) {
              const StagedSample q = stage[g * G + j];
              float ka, kb, va[F2], vb[F2], dv[F2];
              int i0, i1, loc;
              const bool inside = geometry(q, ka, kb, i0, i1, loc);
              cvt_halfs<F2>(dg[j], dv);
              if (lv) {
                if (inside) {
                  lds_feat<F2>(rl + loc, va);
                  lds_feat<F2>(rl + loc + F2, vb);
#pragma unroll
                  for (int f = 0; f < F2; ++f) { va[f] = fmaf(ka, dv[f], va[f]); vb[f] = fmaf(kb, dv[f], vb[f]); }
                  sts_feat<F2>(rl + loc, va);
                  sts_feat<F2>(rl + loc + F2, vb);
                } else {
                  CornerPair<F2> cp;
#pragma unroll
                  for (int f = 0; f < F2; ++f) { cp.a[f] = dv[f]; cp.b[f] = 0.0f; }
                  direct_corner_pair_add<F2>(gkp + goff * F2, i0 + (i1 + c1) * res, cells, ka * scale, kb * scale, cp);
                }
              }
              __syncwarp();   // orders this sample's stores before the next sample's loads (other lanes)
            }
          };
          load_group(0, dA);
#pragma unroll 1
          for (int g = 0; g < NG; g += 2) {
            if (cnt > (g + 1) * G) load_group(g + 1, dB);
            if (cnt > g * G) process_group(g, dA);
            if (g + 2 < NG && cnt > (g + 2) * G) load_group(g + 2, dA);
            if (cnt > (g + 1) * G) process_group(g + 1, dB);
          }
        }
        __syncwarp();   // the stage is rewritten by the next batch
      }
      if constexpr (SCATTER) {
        region_io<F2, REGION_FLUSH>(reg, nullptr, gkp, s_scale, s_res, s_off, s_E, s_base, s_magic, lb, le, ub0, ub1, scale, lane);
      }
      __syncwarp();
    }
  }
}
