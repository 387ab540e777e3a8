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
//   scatter: lane = (level, row parity): the region starts at zero and the warp takes one sample per step.  Of the
//            sample's two window rows each level's lane pair takes the one of its own parity, so a window cell is only
//            ever touched by ONE lane: plain (non-atomic) read-modify-writes of two adjacent corners, ordered by program
//            order alone - no synchronisation between samples, no branch in the loop.  The samples' latent gradients are
//            staged in shared memory by cp.async one batch (32 samples) ahead.  The region is flushed once per task with
//            one vector reduction per non-zero cell: ~8 global reductions per (sample, plane), not 64.  With 8-byte
//            cells and 16 levels the region is bank-interleaved (see grid_binned_kernel): no bank conflicts at all.
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
  // scatter-add with 8-byte cells only: bank-interleaved window layout (see grid_binned_kernel); 0 = packed layout
  int32_t ilv;
  int32_t ilv_rows[2];               // 128-byte region rows of level group 0 (levels 0-7) and 1 (levels 8-15)
  // scatter-add: staging ring of the samples' latent-gradient slices (2 batches of 32 samples, filled by cp.async)
  int32_t dz_extra_floats;           // per-warp floats behind the batch stage (0: the ring lives in unused region rows)
  int32_t dz_base;                   // float offset of the ring from the warp's region
  int32_t dz_stride;                 // bytes from one sample's slice to the next
  int32_t idle_base;                 // float offset of 32 private cell pairs for lanes without a level (n_levels % 16 != 0)
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
  unsigned long long* scan_state;   // [1 + 3 * nt / 1024] ticket counter, then one published word per scan CTA (zeroed per call)
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

#ifndef NVP_SP_UNROLL
#define NVP_SP_UNROLL 4   // rows in flight per thread of the voxel role's scatter-add
#endif
constexpr int kBinThreadsMax = 1024;   // kernel variants are compiled for 512 / 768 / 1024 threads per CTA

// One bucket entry: byte offset of the sample's latent row inside the tile buffer (0xffffffff = padding entry of a
// partial batch), the row's swizzle term ((row & 7) << 4) and the sample's two plane coordinates.
struct __align__(16) StagedSample { uint32_t rowoff, swz; float u0, u1; };
// Per-warp scratch behind the window region: the scatter-add's staged batch (32 x 16 B) or the gather's per-task level
// table (n_levels <= 32 x 32 B).
struct __align__(16) LevelWindow { float scale; int lo0, lo1, base; int E, amax0, amax1, res; };
constexpr int kStageFloats = NVP_MAX_LEVELS * sizeof(LevelWindow) / sizeof(float);
constexpr int kScatterStageFloats = 32 * sizeof(StagedSample) / sizeof(float);   // the scatter-add needs the batch only
__host__ __device__ constexpr int stage_floats(bool scatter) { return scatter ? kScatterStageFloats : kStageFloats; }

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

// Exclusive prefix sums of the bucket sizes (-> offs, cursor) and of the per-bucket task counts, then the task list
// itself; a bucket of c samples becomes ceil(c / chunk) tasks.  One bucket per thread, 1024 buckets per CTA.  The CTAs
// chain through `scan_state`: each takes a ticket (its position in scheduling order, so every predecessor is already
// running), publishes its totals as one 64-bit word (valid bit | samples | tasks) and adds up its predecessors' words.
__global__ void __launch_bounds__(1024) grid_bin_scan_kernel(const BinArgs a) {
  __shared__ int s_wc[32], s_wt[32], s_block, s_cpre, s_tpre;
  const int m = 3 * a.bt.nt, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_block = static_cast<int>(atomicAdd(reinterpret_cast<unsigned int*>(a.scan_state), 1u));
  __syncthreads();
  const int b = s_block, i = b * 1024 + static_cast<int>(threadIdx.x);
  const int c = i < m ? __ldg(a.cnt + i) : 0;
  const int t = c == 0 ? 0 : (c <= a.bt.chunk ? 1 : (c + a.bt.chunk - 1) / a.bt.chunk);
  int ic = c, it = t;   // inclusive scans: warp, then across the warps
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int vc = __shfl_up_sync(0xffffffffu, ic, d), vt = __shfl_up_sync(0xffffffffu, it, d);
    if (lane >= d) { ic += vc; it += vt; }
  }
  if (lane == 31) { s_wc[warp] = ic; s_wt[warp] = it; }
  __syncthreads();
  if (warp == 0) {
    int wc = s_wc[lane], wt = s_wt[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int vc = __shfl_up_sync(0xffffffffu, wc, d), vt = __shfl_up_sync(0xffffffffu, wt, d);
      if (lane >= d) { wc += vc; wt += vt; }
    }
    if (lane == 31) {
      __threadfence();
      *reinterpret_cast<volatile unsigned long long*>(a.scan_state + 1 + b) =
          (1ull << 63) | (static_cast<unsigned long long>(wc) << 32) | static_cast<unsigned long long>(static_cast<unsigned int>(wt));
    }
    s_wc[lane] = wc; s_wt[lane] = wt;   // inclusive over the warps
    // predecessors' totals
    int pc = 0, pt = 0;
    for (int k = lane; k < b; k += 32) {
      unsigned long long w;
      do { w = *reinterpret_cast<volatile unsigned long long*>(a.scan_state + 1 + k); } while (!(w >> 63));
      pc += static_cast<int>((w >> 32) & 0x7fffffffu);
      pt += static_cast<int>(w & 0xffffffffu);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { pc += __shfl_xor_sync(0xffffffffu, pc, d); pt += __shfl_xor_sync(0xffffffffu, pt, d); }
    if (lane == 0) { s_cpre = pc; s_tpre = pt; }
  }
  __syncthreads();
  const int wbase_c = warp ? s_wc[warp - 1] : 0, wbase_t = warp ? s_wt[warp - 1] : 0;
  const int off = s_cpre + wbase_c + ic - c;
  int to = s_tpre + wbase_t + it - t;
  if (i < m) {
    a.offs[i] = off; a.cursor[i] = off;
    for (int bb = 0; bb < c; bb += a.bt.chunk) a.tasks[to++] = make_int2(i, off + bb);
  }
  if (i == m - 1) { a.offs[m] = off + c; a.n_tasks[0] = to; }
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

// The same through a 32-bit shared-window address (keeps the address arithmetic of the scatter-add's hot loop to one add).
template <int F>
__device__ __forceinline__ void lds_feat_s(uint32_t addr, float (&v)[F]) {
  if constexpr (F == 1) {
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(addr) : "memory");
  } else if constexpr (F == 2) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(addr) : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < F; i += 4)
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[i]), "=f"(v[i + 1]), "=f"(v[i + 2]), "=f"(v[i + 3])
                   : "r"(addr + i * 4) : "memory");
  }
}
template <int F>
__device__ __forceinline__ void sts_feat_s(uint32_t addr, const float (&v)[F]) {
  if constexpr (F == 1) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v[0]) : "memory");
  } else if constexpr (F == 2) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v[0]), "f"(v[1]) : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < F; i += 4)
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr + i * 4), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]),
                   "f"(v[i + 3]) : "memory");
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

// Interleaved window layout of the scatter-add (ILV; 8-byte cells, 16 levels).  In the packed layout a lane's cell sits
// at an arbitrary bank and the 32 lanes of a sample's read-modify-write collide (measured on config S: 103 M shared
// wavefronts per step, 54 M of them bank conflicts).  Here every (level, row parity) pair - i.e. every lane - owns one fixed
// 8-byte bank pair: a 128-byte region row holds slot s of the 16 pairs of one level group (levels 0-7 | 8-15), and the
// window cell (aa, row) of level l lives in slot (row >> 1) * E_l + aa of pair (l & 7, row & 1).  The 16 lanes of a
// half-warp always hit 16 different bank pairs: every 64-bit access is conflict-free (57 M wavefronts, 5 M conflicts, all
// in the flush).  The price is a region as tall as the largest window of each group (13.25 KiB instead of 5 KiB for config
// S: 16 windows per SM instead of 22); the rows that the group's small levels leave unused hold the gradient staging ring,
// and the flush walks a per-CTA table of the populated (slot, pair) entries instead of the windows.
constexpr int kIlvOffBits = 14, kIlvLevelBits = 4, kIlvCoordBits = 6;

template <int F2, bool SCATTER, int THREADS, bool ILV = false>
__global__ void __launch_bounds__(THREADS, 1) grid_binned_kernel(const BinArgs a) {
  static_assert(!ILV || (SCATTER && F2 == 2), "the interleaved layout is the scatter-add's, for 8-byte cells");
  extern __shared__ __align__(16) float s_region[];
  __shared__ int s_nitems;
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
  const int region_floats = ILV ? (a.bt.ilv_rows[0] + a.bt.ilv_rows[1]) * 32 : (a.bt.base[L] * F2 + 3) & ~3;
  const int warp_floats = region_floats + stage_floats(SCATTER) + (SCATTER ? a.bt.dz_extra_floats : 0);
  uint32_t* s_items = reinterpret_cast<uint32_t*>(s_region + static_cast<size_t>(nwarps) * warp_floats);
  if constexpr (ILV) {
    if (warp == 0) {   // flush table: the populated entries of a region, row-major (consecutive entries = different banks)
      const int rows0 = a.bt.ilv_rows[0], rows = rows0 + a.bt.ilv_rows[1];
      int n = 0;
      for (int c0 = 0; c0 < rows * 16; c0 += 32) {
        const int c = c0 + lane, row = c >> 4, p = c & 15;
        const int g = row >= rows0 ? 1 : 0, slot = row - (g ? rows0 : 0), l = g * 8 + (p >> 1);
        bool valid = false;
        uint32_t e = 0;
        if (row < rows && l < L) {
          const int E = s_E[l], q = slot / E, aa = slot - q * E, wrow = 2 * q + (p & 1);
          valid = wrow < E;
          e = static_cast<uint32_t>(row * 32 + p * 2) | static_cast<uint32_t>(l) << kIlvOffBits |
              static_cast<uint32_t>(aa) << (kIlvOffBits + kIlvLevelBits) |
              static_cast<uint32_t>(wrow) << (kIlvOffBits + kIlvLevelBits + kIlvCoordBits);
        }
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        if (valid) s_items[n + __popc(m & ((1u << lane) - 1u))] = e;
        n += __popc(m);
      }
      if (lane == 0) s_nitems = n;
    }
    __syncthreads();
  }
  if (warp >= nwarps) {
    if constexpr (F2 <= 4) {   // (the role is only enabled when the 3-D grid has F2 features per voxel)
      const int64_t gtid = (static_cast<int64_t>(blockIdx.x) * a.sp_warps + (warp - nwarps)) * 32 + lane;
      const int64_t gstride = static_cast<int64_t>(gridDim.x) * a.sp_warps * 32;
      if constexpr (SCATTER) {
        if (a.sp.gsparse != nullptr) {
          const float sscale = a.sp.scale_ptr ? a.sp.scale * __ldg(a.sp.scale_ptr) : a.sp.scale;
          // four rows' loads in flight per thread before their reductions (few warps run this role: latency-bound)
          constexpr int U = NVP_SP_UNROLL;
          const int64_t total = 3 * a.sp.n;
          int64_t idx = gtid;
          for (; idx + (U - 1) * gstride < total; idx += U * gstride) {
            SparseRowIn<F2> in[U];
#pragma unroll
            for (int u = 0; u < U; ++u) in[u] = sparse_row_load<F2>(a.sp, a.sp_col0, idx + u * gstride);
#pragma unroll
            for (int u = 0; u < U; ++u) sparse_row_commit<F2>(a.sp, sscale, idx + u * gstride, in[u]);
          }
          for (; idx < total; idx += gstride) sparse_scatter_row<F2>(a.sp, a.sp_col0, sscale, idx);
        }
      } else {
#pragma unroll 2
        for (int64_t s = gtid; s < a.sp.n_pad; s += gstride) sparse_pad_sample<F2>(a.sp, a.sp_col0, s);
      }
    }
    return;
  }
  float* reg = s_region + static_cast<size_t>(warp) * warp_floats;
  StagedSample* stage = reinterpret_cast<StagedSample*>(reg + region_floats);
  const int n_tasks = __ldg(a.n_tasks);
  const float inv_tb = 1.0f / static_cast<float>(a.bt.tb);
  const int pw = L * F2;
  float scale = 1.0f;
  if constexpr (SCATTER) scale = a.scale_ptr ? a.scale * __ldg(a.scale_ptr) : a.scale;
  const int par = lane & 1;   // scatter-add: the window-row parity this lane owns

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
      // ---- scatter-add: lane = (level, row parity); one sample per step of the warp
      const int le = min(L, lb + 16);
      const bool lv = lb + (lane >> 1) < L;
      const int l = lv ? lb + (lane >> 1) : L - 1;
      // per-lane constants of this task: its level's window
      const float sc = s_scale[l];
      const int res = s_res[l], cells = res * res, E = s_E[l];
      const size_t goff = static_cast<size_t>(s_off[l]);
      const int lo0 = static_cast<int>(floorf(fmaf(sc, ub0, 0.5f))), lo1 = static_cast<int>(floorf(fmaf(sc, ub1, 0.5f)));
      const unsigned amax0 = static_cast<unsigned>(min(E - 2, res - 1 - lo0)), amax1 = static_cast<unsigned>(min(E - 2, res - 1 - lo1));
      // Row ownership: of a sample's two window rows bb, bb + 1 the lane takes the one whose parity is its own.  A window
      // cell is then only ever touched by one lane, so consecutive samples need no ordering between lanes: each lane's
      // read-modify-writes are ordered by program order alone.
      // This lane's window base: packed = the level's window; interleaved = the bank pair of (level, parity).
      float* rl = ILV ? reg + ((lane >> 4) ? a.bt.ilv_rows[0] * 32 : 0) + (((lane >> 1) & 7) << 2) + (par << 1)
                      : reg + static_cast<size_t>(s_base[l]) * F2;
      constexpr int kNextCorner = ILV ? 32 : F2;   // floats from the cell (aa, row) to (aa + 1, row)
      // The samples' latent gradients (this plane's slice of the row: NC 16-byte chunks) are staged in shared memory one
      // batch ahead by cp.async: the copies of batch i + 1 are in flight while batch i is processed, so the warp never
      // waits for them (in registers only 16 samples could be kept in flight, and the kernel was bound by that latency).
      const int NC = pw >> 3, chunk0 = plane * NC;
      const uint32_t nc_magic = 65536u / static_cast<uint32_t>(NC) + 1u;   // idx / NC for idx < 32 * NC <= 512
      uint8_t* dzs = reinterpret_cast<uint8_t*>(reg + a.bt.dz_base);
      const uint32_t dz_stride = static_cast<uint32_t>(a.bt.dz_stride), dz_buf = 32u * dz_stride;
      const uint8_t* my_dz = dzs + l * F2 * 2;   // this lane's level inside a staged slice
      auto issue_dz = [&](const uint4& r, uint32_t buf) {   // r = this lane's record of that batch (lane = sample)
        for (int idx = lane; idx < 32 * NC; idx += 32) {
          const int smp = static_cast<int>((static_cast<uint32_t>(idx) * nc_magic) >> 16), q = idx - smp * NC;
          const uint32_t rowoff = __shfl_sync(0xffffffffu, r.x, smp), swz = __shfl_sync(0xffffffffu, r.y, smp);
          const int c = chunk0 + q;
          const bool real = rowoff != 0xffffffffu;
          const uint8_t* src = real ? a.z16t + static_cast<size_t>(c >> 3) * tc::kPanelBytes + rowoff +
                                          ((static_cast<uint32_t>(c & 7) << 4) ^ swz)
                                    : zero_src;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(dzs + buf * dz_buf + smp * dz_stride + q * 16)),
                       "l"(src), "r"(real ? 16u : 0u) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };

      uint4 rec = beg + lane < end ? __ldg(recs + beg + lane) : pad_rec;
      uint4 rec_next = beg + 32 + lane < end ? __ldg(recs + beg + 32 + lane) : pad_rec;
      {
        float4* r4 = reinterpret_cast<float4*>(reg);
        for (int i = lane; i < region_floats / 4; i += 32) r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();
      issue_dz(rec, 0u);

      // One sample, prepared ahead of its turn: byte offset of the lane's first corner from its window base, the two
      // corner weights and the gradient bits.  The hot loop has no branch: a sample outside the window (coordinates outside
      // [0,1] - rare) and an idle lane (level >= L) add zero to a cell that only this lane touches, and the batch is
      // revisited afterwards by the direct path if any lane saw an outside sample.
      constexpr uint32_t kNextCornerBytes = kNextCorner * 4;
      const uint32_t rl_s = lv ? tc::smem_u32(rl) : tc::smem_u32(reg + a.bt.idle_base + lane * (2 * F2));
      const int own_loc = lv ? (ILV ? 0 : par * E * F2 * 4) : 0;   // cell (0, parity row): owned by this lane
      bool saw_outside = false;
      struct Prepared { float ka, kb; int loc; RawHalfs<F2> d; };
      auto prepare = [&](int j, uint32_t buf) -> Prepared {
        const StagedSample q = stage[j];
        Prepared g;
        const float p0 = fmaf(sc, q.u0, 0.5f), p1 = fmaf(sc, q.u1, 0.5f);
        const float f0 = floorf(p0), f1 = floorf(p1);
        const float w0 = p0 - f0, w1 = p1 - f1;
        const int aa = static_cast<int>(f0) - lo0, bb = static_cast<int>(f1) - lo1;
        const int c1 = (bb ^ par) & 1, wrow = bb + c1;     // the lane's row of this sample
        const float wr = c1 ? w1 : 1.0f - w1;
        g.ka = (1.0f - w0) * wr; g.kb = w0 * wr;
        g.loc = ILV ? ((wrow >> 1) * E + aa) << 7 : (wrow * E + aa) * (F2 * 4);
        g.d = lds_halfs_raw<F2>(my_dz + buf * dz_buf + j * dz_stride);
        const bool inside = static_cast<unsigned>(aa) <= amax0 && static_cast<unsigned>(bb) <= amax1;
        saw_outside |= !inside;
        if (!(inside && lv)) {
          g.loc = own_loc;
#pragma unroll
          for (int w = 0; w < (F2 + 1) / 2; ++w) g.d.w[w] = 0u;
        }
        return g;
      };

      uint32_t buf = 0;
      for (int i = beg; i < end; i += 32, buf ^= 1u) {
        reinterpret_cast<uint4*>(stage)[lane] = rec;
        const uint4 rec_next2 = i + 64 + lane < end ? __ldg(recs + i + 64 + lane) : pad_rec;   // two batches ahead
        if (i + 32 < end) issue_dz(rec_next, buf ^ 1u);   // (that buffer was drained by the previous batch)
        else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // this batch's copies have landed
        __syncwarp();
        const int cnt = min(32, end - i);
        Prepared cur = prepare(0, buf);
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
          const Prepared nxt = prepare(min(j + 1, 31), buf);   // independent of the read-modify-write below
          float va[F2], vb[F2], dv[F2];
          cvt_halfs<F2>(cur.d, dv);
          const uint32_t addr = rl_s + static_cast<uint32_t>(cur.loc);
          lds_feat_s<F2>(addr, va);
          lds_feat_s<F2>(addr + kNextCornerBytes, vb);
#pragma unroll
          for (int f = 0; f < F2; ++f) { va[f] = fmaf(cur.ka, dv[f], va[f]); vb[f] = fmaf(cur.kb, dv[f], vb[f]); }
          sts_feat_s<F2>(addr, va);
          sts_feat_s<F2>(addr + kNextCornerBytes, vb);
          cur = nxt;
        }
        if (__any_sync(0xffffffffu, saw_outside && lv)) {
          // direct global reductions for the samples outside this task's window
          for (int j = 0; j < cnt; ++j) {
            const StagedSample q = stage[j];
            const float p0 = fmaf(sc, q.u0, 0.5f), p1 = fmaf(sc, q.u1, 0.5f);
            const float f0 = floorf(p0), f1 = floorf(p1);
            const int i0 = static_cast<int>(f0), i1 = static_cast<int>(f1);
            const int aa = i0 - lo0, bb = i1 - lo1;
            if (!lv || (static_cast<unsigned>(aa) <= amax0 && static_cast<unsigned>(bb) <= amax1)) continue;
            const float w0 = p0 - f0, w1 = p1 - f1;
            const int c1 = (bb ^ par) & 1;
            const float wr = c1 ? w1 : 1.0f - w1;
            CornerPair<F2> cp;
            cvt_halfs<F2>(lds_halfs_raw<F2>(my_dz + buf * dz_buf + j * dz_stride), cp.a);
#pragma unroll
            for (int f = 0; f < F2; ++f) cp.b[f] = 0.0f;
            direct_corner_pair_add<F2>(gkp + goff * F2, i0 + (i1 + c1) * res, cells, (1.0f - w0) * wr * scale, w0 * wr * scale, cp);
          }
          saw_outside = false;
        }
        __syncwarp();   // the stage and this staging buffer are rewritten by the following batches
        rec = rec_next; rec_next = rec_next2;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if constexpr (ILV) {
        const int n_items = s_nitems;
        for (int it = lane; it < n_items; it += 32) {
          const uint32_t e = s_items[it];
          float v[F2];
          lds_feat<F2>(reg + (e & ((1u << kIlvOffBits) - 1u)), v);
          if (v[0] == 0.0f && v[1] == 0.0f) continue;
          const int fl = (e >> kIlvOffBits) & ((1 << kIlvLevelBits) - 1);
          const int aa = (e >> (kIlvOffBits + kIlvLevelBits)) & ((1 << kIlvCoordBits) - 1);
          const int wrow = (e >> (kIlvOffBits + kIlvLevelBits + kIlvCoordBits)) & ((1 << kIlvCoordBits) - 1);
          const float fsc = s_scale[fl];
          const int fres = s_res[fl];
          const int g0 = static_cast<int>(floorf(fmaf(fsc, ub0, 0.5f))) + aa, g1 = static_cast<int>(floorf(fmaf(fsc, ub1, 0.5f))) + wrow;
          if (g0 > fres || g1 > fres) continue;   // (res, j) and (i, res) are the aliased "virtual" cells
#pragma unroll
          for (int f = 0; f < F2; ++f) v[f] *= scale;
          red_feat<F2>(gkp + (static_cast<size_t>(s_off[fl]) + wrap_cell(g0 + g1 * fres, fres * fres)) * F2, v);
        }
      } else if constexpr (SCATTER) {
        region_io<F2, REGION_FLUSH>(reg, nullptr, gkp, s_scale, s_res, s_off, s_E, s_base, s_magic, lb, le, ub0, ub1, scale, lane);
      }
      __syncwarp();
    }
  }
}
