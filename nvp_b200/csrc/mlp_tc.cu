// Tensor-core path (mode NVP_MODE_TC_F16): the modulated SIREN as fused tcgen05 kernels.
//
// Reference semantics: modulation.py:83-92 (SirenNet.forward), :112-121 (Modulator.forward),
// modules.py:78-82, loss_functions.py:3, training.py:47-48,74.  Arithmetic: fp16 operands, fp32
// accumulation in TMEM, fp32 epilogues (bias, LeakyReLU, range-reduced sin/cos, gating, head).
//
// One CTA per SM, persistent over 128-sample tiles.  Warp roles:
//   warp 0      TMA producer   : cp.async.bulk of the latent tile and of the weight panels (ring)
//   warp 1      MMA issuer     : tcgen05.mma (one thread), accumulators in TMEM, tcgen05.commit
//   warps 2..9  epilogue       : tcgen05.ld -> fp32 math -> fp16 A-operand tiles for the next layer
//                                written straight into shared memory in the UMMA swizzled layout
// Every operand (activations and weights) uses the panel format of tc_common.cuh, in HBM as well as in
// shared memory, so all global->shared traffic is plain bulk TMA and the same stored activation tile
// serves the forward, dgrad (K-major) and wgrad (MN-major) GEMMs.
#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

namespace nvp {
namespace {
using namespace tc;

constexpr int H = kHidden;          // 128
constexpr int kTile = 128;          // samples per tile (UMMA M)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;  // 320
constexpr int kMaxPack = 72;

// ------------------------------------------------------------------------------------------
// Weight packing: fp32 nn.Linear weights -> fp16 panels in MMA stream order (runs every call,
// ~1 MB of traffic).
// ------------------------------------------------------------------------------------------
struct PackPanel {
  const float* src;
  int ld;          // row pitch of src (floats)
  int transpose;   // 0: panel(r,c) = src[(r0+r)*ld + c0+c] ; 1: panel(r,c) = src[(c0+c)*ld + r0+r]
  int r0, c0;
  int rvalid, cvalid;  // elements outside are zero
  int rows;        // panel rows (128 or ZP)
  uint32_t dst_off;
};
struct PackArgs {
  PackPanel p[kMaxPack];
  int n;
  uint8_t* dst;
};

__global__ void __launch_bounds__(256) pack_weights_kernel(const PackArgs a) {
  const PackPanel& pp = a.p[blockIdx.x];
  uint8_t* dst = a.dst + pp.dst_off;
  for (int e = threadIdx.x; e < pp.rows * 64; e += blockDim.x) {
    int r, c;
    if (pp.transpose) { r = e % pp.rows; c = e / pp.rows; } else { r = e >> 6; c = e & 63; }
    float v = 0.0f;
    if (r < pp.rvalid && c < pp.cvalid)
      v = pp.transpose ? __ldg(pp.src + static_cast<size_t>(pp.c0 + c) * pp.ld + pp.r0 + r)
                       : __ldg(pp.src + static_cast<size_t>(pp.r0 + r) * pp.ld + pp.c0 + c);
    *reinterpret_cast<__half*>(dst + panel_offset(r, c)) = __float2half_rn(v);
  }
}

struct Dims {
  int Z, ZP, KZ;
  int npf;            // forward weight panels per tile
  size_t fwd_bytes;   // bytes of the forward panel stream
};
Dims make_dims(const nvp_desc* d) {
  Dims m;
  m.Z = latent_dim(d);
  m.ZP = round_up(m.Z + 1, 64);  // +1: spare column carrying the constant 1 (bias gradients)
  m.KZ = m.ZP / 64;
  m.npf = 8 + 3 * m.KZ;
  m.fwd_bytes = static_cast<size_t>(m.npf) * kPanelBytes;
  return m;
}

void add_panel(PackArgs& a, const float* src, int ld, int transpose, int r0, int c0, int rvalid, int cvalid, int rows,
               uint32_t& off) {
  PackPanel& p = a.p[a.n++];
  p.src = src; p.ld = ld; p.transpose = transpose; p.r0 = r0; p.c0 = c0;
  p.rvalid = std::max(0, std::min(rvalid, rows)); p.cvalid = std::max(0, std::min(cvalid, 64));
  p.rows = rows; p.dst_off = off;
  off += static_cast<uint32_t>(rows) * 128u;
}

// Forward stream: [W0z] | [W1h][W1z][Ws1] | [W2h][W2z][Ws2], each as 64-wide K panels of a [128 out x K] matrix.
int pack_forward_weights(const nvp_desc* d, const nvp_params* p, uint8_t* dst, cudaStream_t st) {
  const Dims m = make_dims(d);
  PackArgs a{};
  a.dst = dst;
  uint32_t off = 0;
  for (int q = 0; q < m.KZ; ++q) add_panel(a, p->mod_w[0], m.Z, 0, 0, 64 * q, H, m.Z - 64 * q, H, off);
  for (int i = 1; i < 3; ++i) {
    for (int q = 0; q < 2; ++q) add_panel(a, p->mod_w[i], H + m.Z, 0, 0, 64 * q, H, 64, H, off);
    for (int q = 0; q < m.KZ; ++q) add_panel(a, p->mod_w[i], H + m.Z, 0, 0, H + 64 * q, H, m.Z - 64 * q, H, off);
    for (int q = 0; q < 2; ++q) add_panel(a, p->siren_w[i], H, 0, 0, 64 * q, H, 64, H, off);
  }
  pack_weights_kernel<<<a.n, 256, 0, st>>>(a);
  NVP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Shared epilogue helpers
// ------------------------------------------------------------------------------------------
// sin/cos with two-constant Cody-Waite reduction to [-pi, pi] then the SFU approximation
// (abs err ~4e-7 after reduction; arguments reach |30*(w*t+b)| ~ 60, modulation.py:25,69).
__device__ __forceinline__ float reduce_2pi(float x) {
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, x);
  return fmaf(k, 1.7484556e-7f, r);
}
__device__ __forceinline__ float fast_sin(float x) { return __sinf(reduce_2pi(x)); }
__device__ __forceinline__ void fast_sincos(float x, float& s, float& c) { __sincosf(reduce_2pi(x), &s, &c); }
__device__ __forceinline__ float lrelu(float x) { return fmaxf(x, 0.01f * x); }

// Store 32 consecutive columns (starting at panel column c0, multiple of 32) of row r into a panel as fp16.
__device__ __forceinline__ void store_row32(uint8_t* panel, int r, int c0, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 q;
    q.x = pack_half2(v[8 * j + 0], v[8 * j + 1]);
    q.y = pack_half2(v[8 * j + 2], v[8 * j + 3]);
    q.z = pack_half2(v[8 * j + 4], v[8 * j + 5]);
    q.w = pack_half2(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(panel + panel_chunk_offset(r, (c0 >> 3) + j)) = q;
  }
}

// Stash slots per tile (train mode), each 2 panels = 32 KiB.
enum { SL_H0 = 0, SL_A0, SL_H1, SL_A1, SL_S1, SL_C1, SL_H2, SL_S2, SL_C2, SL_COUNT };

struct FwdArgs {
  const uint8_t* wpk;     // forward weight panel stream
  const uint8_t* z16t;    // latent tiles
  const float* tau;       // [n]
  const float* mod_b[3];
  const float* siren_b[3];
  const float* siren_w0;  // [128] (net.layers.0.weight [128,1])
  const float* last_w;    // [3,128]
  const float* last_b;    // [3]
  float w0;
  float* rgb;             // [n,3]
  uint8_t* stash;         // train: [tile][SL_COUNT][2 panels]
  int64_t n;
  int n_tiles, KZ, npf, nstage;
};

struct FwdSmem {  // offsets into dynamic smem (1024-B aligned base)
  uint32_t z, h, a, ring, consts, rgbx, bars, total;
};
__host__ __device__ inline FwdSmem fwd_smem_layout(int KZ, int nstage) {
  FwdSmem s;
  uint32_t o = 0;
  s.z = o; o += KZ * kPanelBytes;
  s.h = o; o += 2 * kPanelBytes;
  s.a = o; o += 2 * kPanelBytes;
  s.ring = o; o += nstage * kPanelBytes;
  s.consts = o; o += (10 * H + 4) * 4;      // bm[3][H] bs[3][H] ws0[H] wl[3][H] bl[3]
  s.rgbx = o; o += H * 3 * 4;               // partial rgb of the upper column half
  s.bars = o; o += 64 * 8;
  s.total = o;
  return s;
}

template <bool TRAIN>
__global__ void __launch_bounds__(kThreads, 1) mlp_forward_kernel(const FwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const FwdSmem L = fwd_smem_layout(a.KZ, a.nstage);
  uint8_t* zbuf = smem + L.z;
  uint8_t* hbuf = smem + L.h;
  uint8_t* abuf = smem + L.a;
  uint8_t* ring = smem + L.ring;
  float* cst = reinterpret_cast<float*>(smem + L.consts);
  float* s_bm = cst;            // [3][H]
  float* s_bs = cst + 3 * H;    // [3][H]
  float* s_ws0 = cst + 6 * H;   // [H]
  float* s_wl = cst + 7 * H;    // [3][H]
  float* s_bl = cst + 10 * H;   // [3]
  float* s_rgbx = reinterpret_cast<float*>(smem + L.rgbx);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* wfull = bars;                 // [nstage]
  uint64_t* wempty = bars + 16;           // [nstage]
  uint64_t* zfull = bars + 32;
  uint64_t* zempty = bars + 33;
  uint64_t* acc_full = bars + 34;
  uint64_t* epi_done = bars + 35;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 3 * H; i += kThreads) {
    s_bm[i] = __ldg(a.mod_b[i / H] + (i % H));
    s_bs[i] = __ldg(a.siren_b[i / H] + (i % H));
    s_wl[i] = __ldg(a.last_w + i);
  }
  for (int i = tid; i < H; i += kThreads) s_ws0[i] = __ldg(a.siren_w0 + i);
  if (tid < 3) s_bl[tid] = __ldg(a.last_b + tid);
  if (tid == 0) {
    for (int i = 0; i < a.nstage; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    mbar_init(zfull, 1); mbar_init(zempty, 1); mbar_init(acc_full, 1); mbar_init(epi_done, kEpiWarps);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&s_tmem, 256); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t acc_m = tmem, acc_s = tmem + 128;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t g = 0, it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(zempty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(zfull, a.KZ * kPanelBytes);
        bulk_g2s(zbuf, a.z16t + static_cast<size_t>(tile) * a.KZ * kPanelBytes, a.KZ * kPanelBytes, zfull);
        for (int i = 0; i < a.npf; ++i, ++g) {
          const uint32_t st = g % a.nstage, ph = (g / a.nstage) & 1;
          mbar_wait(&wempty[st], ph ^ 1);
          mbar_arrive_expect_tx(&wfull[st], kPanelBytes);
          bulk_g2s(ring + st * kPanelBytes, a.wpk + static_cast<size_t>(i) * kPanelBytes, kPanelBytes, &wfull[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(kTile, H, false, false);
      uint32_t g = 0, it = 0, n_epi = 0;
      auto gemm_panel = [&](uint32_t a_panel_addr, uint32_t acc, bool& first) {
        const uint32_t st = g % a.nstage, ph = (g / a.nstage) & 1;
        mbar_wait(&wfull[st], ph);
        tcgen05_fence_after();
        const uint32_t b_addr = smem_u32(ring + st * kPanelBytes);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          umma_f16_ss(acc, umma_desc_kmajor(a_panel_addr, kk), umma_desc_kmajor(b_addr, kk), idesc, first ? 0u : 1u);
          first = false;
        }
        umma_commit(&wempty[st]);
        ++g;
      };
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(zfull, it & 1);
        for (int step = 0; step < 3; ++step) {
          if (!(it == 0 && step == 0)) { mbar_wait(epi_done, n_epi & 1); ++n_epi; }
          tcgen05_fence_after();
          bool first = true;
          if (step > 0)
            for (int q = 0; q < 2; ++q) gemm_panel(smem_u32(hbuf + q * kPanelBytes), acc_m, first);
          for (int q = 0; q < a.KZ; ++q) gemm_panel(smem_u32(zbuf + q * kPanelBytes), acc_m, first);
          if (step == 2) umma_commit(zempty);
          if (step > 0) {
            first = true;
            for (int q = 0; q < 2; ++q) gemm_panel(smem_u32(abuf + q * kPanelBytes), acc_s, first);
          }
          umma_commit(acc_full);
        }
      }
    }
  } else {
    // ================= epilogue warps =================
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;          // which 64-column half / output panel
    const int r = quarter * 32 + lane;         // row inside the tile
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t n_acc = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int64_t s = static_cast<int64_t>(tile) * kTile + r;
      const bool valid = s < a.n;
      const float tau = valid ? __ldg(a.tau + s) : 0.0f;
      uint8_t* st_base = TRAIN ? a.stash + (static_cast<size_t>(tile) * SL_COUNT) * 2 * kPanelBytes : nullptr;
      auto stash_ptr = [&](int slot) { return st_base + (static_cast<size_t>(slot) * 2 + half) * kPanelBytes; };
      float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
      for (int step = 0; step < 3; ++step) {
        mbar_wait(acc_full, n_acc & 1); ++n_acc;
        tcgen05_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int pc = cc * 32;            // column inside my panel
          const int col = half * 64 + pc;    // column inside the layer
          uint32_t vm[32];
          tmem_ld32(acc_m + lane_base + col, vm);
          float hv[32], av[32];
          if (step == 0) {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              hv[i] = lrelu(__uint_as_float(vm[i]) + s_bm[col + i]);
              const float sp = a.w0 * fmaf(tau, s_ws0[col + i], s_bs[col + i]);
              av[i] = fast_sin(sp) * hv[i];
            }
          } else {
            uint32_t vs[32];
            tmem_ld32(acc_s + lane_base + col, vs);
            tmem_ld_wait();
            float sv[32], cv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              hv[i] = lrelu(__uint_as_float(vm[i]) + s_bm[step * H + col + i]);
              const float sp = __uint_as_float(vs[i]) + s_bs[step * H + col + i];
              if (TRAIN) fast_sincos(sp, sv[i], cv[i]); else sv[i] = fast_sin(sp);
              av[i] = sv[i] * hv[i];
            }
            if (TRAIN) {
              store_row32(stash_ptr(step == 1 ? SL_S1 : SL_S2), r, pc, sv);
              store_row32(stash_ptr(step == 1 ? SL_C1 : SL_C2), r, pc, cv);
            }
          }
          if (step < 2) {
            store_row32(hbuf + half * kPanelBytes, r, pc, hv);
            store_row32(abuf + half * kPanelBytes, r, pc, av);
            if (TRAIN) {
              store_row32(stash_ptr(step == 0 ? SL_H0 : SL_H1), r, pc, hv);
              store_row32(stash_ptr(step == 0 ? SL_A0 : SL_A1), r, pc, av);
            }
          } else {
            if (TRAIN) store_row32(stash_ptr(SL_H2), r, pc, hv);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              rgb0 = fmaf(av[i], s_wl[col + i], rgb0);
              rgb1 = fmaf(av[i], s_wl[H + col + i], rgb1);
              rgb2 = fmaf(av[i], s_wl[2 * H + col + i], rgb2);
            }
          }
        }
        if (step == 2) {
          // combine the two column halves of each row: upper half hands its partial sums over
          if (half == 1) { s_rgbx[r * 3] = rgb0; s_rgbx[r * 3 + 1] = rgb1; s_rgbx[r * 3 + 2] = rgb2; }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          if (half == 0 && valid) {
            a.rgb[s * 3] = rgb0 + s_rgbx[r * 3] + s_bl[0];
            a.rgb[s * 3 + 1] = rgb1 + s_rgbx[r * 3 + 1] + s_bl[1];
            a.rgb[s * 3 + 2] = rgb2 + s_rgbx[r * 3 + 2] + s_bl[2];
          }
        }
        fence_proxy_async_smem();   // operand tiles written with st.shared -> visible to the UMMA (async proxy)
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(epi_done);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

struct TcWorkspace {
  uint8_t* wpk_fwd;
  uint8_t* z16t;
  uint8_t* stash;
  float* rgb;
  size_t total;
};
TcWorkspace carve_tc(const nvp_desc* d, int64_t n, int what, void* base) {
  const Dims m = make_dims(d);
  const int64_t tiles = (n + kTile - 1) / kTile;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? static_cast<uint8_t*>(base) + off : nullptr;
    off += (bytes + 1023) / 1024 * 1024;
    return p;
  };
  TcWorkspace w{};
  w.wpk_fwd = take(m.fwd_bytes);
  w.z16t = take(static_cast<size_t>(tiles) * m.KZ * kPanelBytes);
  w.rgb = reinterpret_cast<float*>(take(static_cast<size_t>(tiles) * kTile * 3 * sizeof(float)));
  if (what == 1) w.stash = take(static_cast<size_t>(tiles) * SL_COUNT * 2 * kPanelBytes);
  w.total = off + 1024;
  return w;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}

int launch_forward(const nvp_desc* d, const nvp_params* p, const TcWorkspace& w, const float* tsteps, int64_t n,
                   float* rgb, bool train, cudaStream_t st) {
  const Dims m = make_dims(d);
  FwdArgs a{};
  a.wpk = w.wpk_fwd; a.z16t = w.z16t; a.tau = tsteps;
  for (int i = 0; i < 3; ++i) { a.mod_b[i] = p->mod_b[i]; a.siren_b[i] = p->siren_b[i]; }
  a.siren_w0 = p->siren_w[0]; a.last_w = p->last_w; a.last_b = p->last_b;
  a.w0 = d->w0_first; a.rgb = rgb; a.stash = w.stash; a.n = n;
  a.n_tiles = static_cast<int>((n + kTile - 1) / kTile);
  a.KZ = m.KZ; a.npf = m.npf;
  const int budget = 227 * 1024 - 2048;
  a.nstage = 0;
  for (int ns = 12; ns >= 2; --ns)
    if (static_cast<int>(fwd_smem_layout(m.KZ, ns).total) <= budget) { a.nstage = ns; break; }
  NVP_CHECK(a.nstage >= 2, "latent too wide for the shared-memory plan of the tensor-core path");
  const size_t smem = fwd_smem_layout(m.KZ, a.nstage).total + 1024;
  const int grid = std::min(a.n_tiles, num_sms());
  if (train) {
    NVP_CUDA(cudaFuncSetAttribute(mlp_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    mlp_forward_kernel<true><<<grid, kThreads, smem, st>>>(a);
  } else {
    NVP_CUDA(cudaFuncSetAttribute(mlp_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    mlp_forward_kernel<false><<<grid, kThreads, smem, st>>>(a);
  }
  NVP_LAUNCH_CHECK();
  return 0;
}

}  // namespace

size_t tc_workspace_bytes(const nvp_desc* d, int64_t n, int what) { return carve_tc(d, std::max<int64_t>(n, 1), what, nullptr).total; }

int tc_forward(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords, const float* tsteps,
               int64_t n, float* out_rgb, void* ws, size_t ws_bytes, cudaStream_t st) {
  NVP_CHECK(ws_bytes >= tc_workspace_bytes(d, n, 0), "workspace too small (see nvp_workspace_bytes)");
  void* base = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
  const TcWorkspace w = carve_tc(d, n, 0, base);
  const Dims m = make_dims(d);
  int rc;
  if ((rc = pack_forward_weights(d, p, w.wpk_fwd, st))) return rc;
  if ((rc = launch_grid_gather(d, tab, p, coords, n, nullptr, 0, w.z16t, m.KZ, st))) return rc;
  return launch_forward(d, p, w, tsteps, n, out_rgb, false, st);
}

int tc_fwd_bwd(const nvp_desc*, const LevelTab&, const nvp_params*, const float*, const float*, const uint8_t*,
               const float*, int64_t, int64_t, const nvp_grads*, float*, float*, void*, size_t, cudaStream_t) {
  set_error("NVP_MODE_TC_F16 backward not built yet");
  return 9;
}

}  // namespace nvp
