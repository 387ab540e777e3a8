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
#include <atomic>
#include <initializer_list>

#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace nvp {
namespace {
using namespace tc;

constexpr int H = kHidden;          // 128
constexpr int kTile = 128;          // samples per tile (UMMA M)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;  // 320

// Optional in-kernel timeline (make TIMELINE=1): clock64 stamps of CTA 0 in its 4th tile, read back with
// nvp_debug_timeline_read / scripts/mlp_timeline.py.  Forward: slot = 16*step + 8*panel + k for the epilogue issuer
// (k: 0 accumulators ready, 1 TMEM loaded, 2 tile stored to smem, 3 after the phase barrier), 64 + 8*step + k for the
// MMA thread (k: 0 step start, 1..2 panel q of h/a available, 3 all MMAs issued).  Compiled out by default.
#ifdef NVP_TIMELINE
__device__ unsigned long long g_timeline[128];
#define NVP_TL(cond, slot) do { if (blockIdx.x == 0 && (cond)) g_timeline[(slot)] = clock64(); } while (0)
#else
#define NVP_TL(cond, slot) do {} while (0)
#endif
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
  // fused kernel: a bias folded into the GEMM through the latent's two constant-1 columns.  Panel column bias_hi_c gets
  // fp16(b[r]), column bias_lo_c gets fp16(b[r] - fp16(b[r])) (-1: not in this panel); the two products sum to b in fp32.
  const float* bias;
  int bias_hi_c, bias_lo_c;
};
struct PackArgs {
  PackPanel p[kMaxPack];
  int n;
  uint8_t* dst;
};

__global__ void __launch_bounds__(256) pack_weights_kernel(const PackArgs a) {
  const PackPanel& pp = a.p[blockIdx.x];
  uint8_t* dst = a.dst + pp.dst_off;
  // blockIdx.y = quarter of the panel: the kernel is latency-bound (dependent load -> 2-byte store per element)
  const int per = pp.rows * 16;
  for (int e = blockIdx.y * per + threadIdx.x; e < (blockIdx.y + 1) * per; e += blockDim.x) {
    int r, c;
    if (pp.transpose) { r = e % pp.rows; c = e / pp.rows; } else { r = e >> 6; c = e & 63; }
    float v = 0.0f;
    if (r < pp.rvalid && c < pp.cvalid)
      v = pp.transpose ? __ldg(pp.src + static_cast<size_t>(pp.c0 + c) * pp.ld + pp.r0 + r)
                       : __ldg(pp.src + static_cast<size_t>(pp.r0 + r) * pp.ld + pp.c0 + c);
    if (pp.bias != nullptr && r < pp.rvalid) {
      const float b = __ldg(pp.bias + pp.r0 + r);
      if (c == pp.bias_hi_c) v = b;
      else if (c == pp.bias_lo_c) v = b - __half2float(__float2half_rn(b));
    }
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
  p.bias = nullptr; p.bias_hi_c = p.bias_lo_c = -1;
  off += static_cast<uint32_t>(rows) * 128u;
}

struct BwdArgs;
int pack_backward_panels(const nvp_desc* d, const nvp_params* p, PackArgs& a, uint32_t base_off, BwdArgs* b);

// Forward stream: [W0z] | [W1z][W1h p0][Ws1 p0][W1h p1][Ws1 p1] | [W2z][W2h p0][Ws2 p0][W2h p1][Ws2 p1], each a
// 64-wide K panel of a [128 out x K] matrix, in the order the MMA warp consumes them.
// When `b` is given the backward stream is packed by the same launch into `bwd_dst`.
int pack_forward_weights(const nvp_desc* d, const nvp_params* p, uint8_t* dst, BwdArgs* b, cudaStream_t st,
                         uint8_t* bwd_dst = nullptr) {
  const Dims m = make_dims(d);
  PackArgs a{};
  a.dst = dst;
  uint32_t off = 0;
  for (int q = 0; q < m.KZ; ++q) add_panel(a, p->mod_w[0], m.Z, 0, 0, 64 * q, H, m.Z - 64 * q, H, off);
  for (int i = 1; i < 3; ++i) {
    for (int q = 0; q < m.KZ; ++q) add_panel(a, p->mod_w[i], H + m.Z, 0, 0, H + 64 * q, H, m.Z - 64 * q, H, off);
    for (int q = 0; q < 2; ++q) {
      add_panel(a, p->mod_w[i], H + m.Z, 0, 0, 64 * q, H, 64, H, off);
      add_panel(a, p->siren_w[i], H, 0, 0, 64 * q, H, 64, H, off);
    }
  }
  if (b != nullptr) pack_backward_panels(d, p, a, static_cast<uint32_t>(bwd_dst - dst), b);
  ScopedKernelTimer timer(K_PACK, st);
  pack_weights_kernel<<<dim3(a.n, 4), 256, 0, st>>>(a);
  NVP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Shared epilogue helpers
// ------------------------------------------------------------------------------------------
// sin/cos with two-constant Cody-Waite reduction to [-pi, pi] then the SFU approximation
// (abs err ~4e-7 after reduction; arguments reach |30*(w*t+b)| ~ 60, modulation.py:25,69).
// XU = true rounds with the magic-number add on the FMA pipe instead of rintf (an FRND on the quarter-rate XU pipe): the
// training epilogues are bound by that pipe (MUFU.SIN + MUFU.COS per element; measured -6 % forward time), while the
// inference forward (one MUFU per element) is bound by the FMA pipe and keeps rintf.
template <bool XU_BOUND>
__device__ __forceinline__ float reduce_2pi(float x) {
  const float y = x * 0.15915494309189535f;
  const float k = XU_BOUND ? __fadd_rn(__fadd_rn(y, 12582912.0f), -12582912.0f) : rintf(y);   // magic: rint for |y| < 2^22
  float r = fmaf(k, -6.2831854820251465f, x);
  return fmaf(k, 1.7484556e-7f, r);
}
template <bool XU_BOUND>
__device__ __forceinline__ float fast_sin(float x) { return __sinf(reduce_2pi<XU_BOUND>(x)); }
__device__ __forceinline__ void fast_sincos(float x, float& s, float& c) { __sincosf(reduce_2pi<true>(x), &s, &c); }
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

template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, uint32_t (&v)[N]) {
  if constexpr (N == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
}
// N (16 or 32) consecutive columns starting at panel column c0 (multiple of N) of row r, as fp16.
template <int N>
__device__ __forceinline__ void store_rown(uint8_t* panel, int r, int c0, const float (&v)[N]) {
#pragma unroll
  for (int j = 0; j < N / 8; ++j) {
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
  uint32_t z, h, a, s, c, ring, consts, rgbx, bars, total;
};
__host__ __device__ inline FwdSmem fwd_smem_layout(int KZ, int nstage, bool stage_sc) {
  const bool train = stage_sc;
  FwdSmem s;
  uint32_t o = 0;
  s.z = o; o += KZ * kPanelBytes;
  s.h = o; o += 2 * kPanelBytes;
  s.a = o; o += 2 * kPanelBytes;
  s.s = o; if (train) o += 2 * kPanelBytes;   // staging of sin / cos tiles for the TMA bulk store
  s.c = o; if (train) o += 2 * kPanelBytes;
  s.ring = o; o += nstage * kPanelBytes;
  s.consts = o; o += (10 * H + 4) * 4;      // bm[3][H] bs[3][H] ws0[H] wl[3][H] bl[3]
  s.rgbx = o; o += 3 * H * 3 * 4;           // partial rgb of the other column slices (up to 3)
  s.bars = o; o += 64 * 8;
  s.total = o;
  return s;
}

// Forward kernel, v2 pipeline.
//   * TMEM holds two accumulator sets (modulator / SIREN pre-activations, 4 x 128 columns): the GEMMs of
//     step i+1 that do not depend on step i's epilogue (latent x W_z) run while that epilogue executes.
//   * The epilogue works panel by panel (64 columns): as soon as panel p of h_i / a_i sits in shared
//     memory the MMA warp issues the K-panel-p GEMMs of step i+1.
//   * Train mode: every stashed activation tile is staged in shared memory in the MMA tile format and
//     leaves through cp.async.bulk (TMA) stores, one 16 KiB panel at a time.
// STAGE_SC: stage the sin/cos stash tiles in shared memory for TMA bulk stores (needs 64 KiB); when the latent is too
// wide for that (config L) they are written with per-lane 16-byte stores instead.
// EW = epilogue warps (8 or 16): each TMEM lane quarter is served by EW/4 warps, every warp owning CW = 256/EW columns of
// the current 64-column panel.  The epilogue's instruction issue bounds this kernel (30 K warp instructions per tile);
// 16 warps give the four schedulers four warps each to interleave instead of two.
template <bool TRAIN, bool STAGE_SC, int EW>
__global__ void __launch_bounds__(64 + EW * 32, 1) mlp_forward_kernel(const FwdArgs a) {
  constexpr int CW = 256 / EW;   // columns per warp inside a 64-column panel
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const FwdSmem L = fwd_smem_layout(a.KZ, a.nstage, TRAIN && STAGE_SC);
  uint8_t* zbuf = smem + L.z;
  uint8_t* hbuf = smem + L.h;
  uint8_t* abuf = smem + L.a;
  uint8_t* sbuf = smem + L.s;
  uint8_t* cbuf = smem + L.c;
  uint8_t* ring = smem + L.ring;
  float* cst = reinterpret_cast<float*>(smem + L.consts);
  float* s_bm = cst;            // [3][H]
  float* s_bs = cst + 3 * H;    // [3][H]   (layer 0 pre-multiplied by w0)
  float* s_ws0 = cst + 6 * H;   // [H]      (pre-multiplied by w0)
  float* s_wl = cst + 7 * H;    // [3][H]
  float* s_bl = cst + 10 * H;   // [3]
  float* s_rgbx = reinterpret_cast<float*>(smem + L.rgbx);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* wfull = bars;                 // [nstage]
  uint64_t* wempty = bars + 16;           // [nstage]
  uint64_t* zfull = bars + 32;
  uint64_t* zempty = bars + 33;
  uint64_t* acc_full = bars + 34;
  uint64_t* panel_done = bars + 36;       // [3 steps][2 panels], one completion per tile each
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 3 * H; i += 64 + EW * 32) {
    s_bm[i] = __ldg(a.mod_b[i / H] + (i % H));
    s_bs[i] = __ldg(a.siren_b[i / H] + (i % H)) * (i < H ? a.w0 : 1.0f);
    s_wl[i] = __ldg(a.last_w + i);
  }
  for (int i = tid; i < H; i += 64 + EW * 32) s_ws0[i] = __ldg(a.siren_w0 + i) * a.w0;
  if (tid < 3) s_bl[tid] = __ldg(a.last_b + tid);
  if (tid == 0) {
    for (int i = 0; i < a.nstage; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    mbar_init(zfull, 1); mbar_init(zempty, 1); mbar_init(acc_full, 1);
    for (int i = 0; i < 6; ++i) mbar_init(&panel_done[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;

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
      uint32_t g = 0, it = 0;
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
        const uint32_t par = it & 1;
        mbar_wait(zfull, par);
        for (int step = 0; step < 3; ++step) {
          NVP_TL(it == 3, 64 + 8 * step);
          const uint32_t b = (3 * it + step) & 1;
          const uint32_t acc_m = tmem + 128 * b, acc_s = tmem + 256 + 128 * b;
          // accumulator set b was last read by the epilogue two steps ago; for step 1 that is the previous
          // tile's last step, whose completion has not been observed yet.
          if (step == 1 && it > 0) { mbar_wait(&panel_done[4], par ^ 1); mbar_wait(&panel_done[5], par ^ 1); }
          tcgen05_fence_after();
          bool first_m = true, first_s = true;
          for (int q = 0; q < a.KZ; ++q) gemm_panel(smem_u32(zbuf + q * kPanelBytes), acc_m, first_m);
          if (step == 2) umma_commit(zempty);
          if (step > 0) {
            for (int q = 0; q < 2; ++q) {
              mbar_wait(&panel_done[(step - 1) * 2 + q], par);
              NVP_TL(it == 3, 64 + 8 * step + 1 + q);
              tcgen05_fence_after();
              gemm_panel(smem_u32(hbuf + q * kPanelBytes), acc_m, first_m);
              gemm_panel(smem_u32(abuf + q * kPanelBytes), acc_s, first_s);
            }
          }
          umma_commit(acc_full);
          NVP_TL(it == 3, 64 + 8 * step + 3);
        }
      }
    }
  } else {
    // ================= epilogue warps =================
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int sub = (warp - 2) >> 2;           // which CW-column slice of the current 64-column panel
    const int r = quarter * 32 + lane;         // row inside the tile
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const bool issuer = (tid == 64);
    uint32_t n_acc = 0, it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int64_t s = static_cast<int64_t>(tile) * kTile + r;
      const bool valid = s < a.n;
      const float tau = valid ? __ldg(a.tau + s) : 0.0f;
      uint8_t* st_base = TRAIN ? a.stash + (static_cast<size_t>(tile) * SL_COUNT) * 2 * kPanelBytes : nullptr;
      float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
      for (int step = 0; step < 3; ++step) {
        const uint32_t b = (3 * it + step) & 1;
        const uint32_t acc_m = tmem + 128 * b, acc_s = tmem + 256 + 128 * b;
        mbar_wait(acc_full, n_acc & 1); ++n_acc;
        tcgen05_fence_after();
        NVP_TL(issuer && it == 3, 16 * step);
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
          const int pc = sub * CW;           // column inside panel p
          const int col = p * 64 + pc;       // column inside the layer
          if (TRAIN) {
            // staging buffers of this panel were handed to the TMA two phases ago; at most the newest group may be pending
            if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
          }
          NVP_TL(issuer && it == 3 && p == 1, 16 * step + 8);
          uint32_t vm[CW];
          tmem_ldn<CW>(acc_m + lane_base + col, vm);
          float hv[CW], av[CW];
          if (step == 0) {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < CW; ++i) {
              hv[i] = lrelu(__uint_as_float(vm[i]) + s_bm[col + i]);
              av[i] = fast_sin<TRAIN>(fmaf(tau, s_ws0[col + i], s_bs[col + i])) * hv[i];
            }
          } else {
            uint32_t vs[CW];
            tmem_ldn<CW>(acc_s + lane_base + col, vs);
            tmem_ld_wait();
            float sv[CW], cv[CW];
#pragma unroll
            for (int i = 0; i < CW; ++i) {
              hv[i] = lrelu(__uint_as_float(vm[i]) + s_bm[step * H + col + i]);
              const float sp = __uint_as_float(vs[i]) + s_bs[step * H + col + i];
              if (TRAIN) fast_sincos(sp, sv[i], cv[i]); else sv[i] = fast_sin<false>(sp);
              av[i] = sv[i] * hv[i];
            }
            if (TRAIN && STAGE_SC) {
              store_rown<CW>(sbuf + p * kPanelBytes, r, pc, sv);
              store_rown<CW>(cbuf + p * kPanelBytes, r, pc, cv);
            } else if (TRAIN) {
              store_rown<CW>(st_base + (static_cast<size_t>(step == 1 ? SL_S1 : SL_S2) * 2 + p) * kPanelBytes, r, pc, sv);
              store_rown<CW>(st_base + (static_cast<size_t>(step == 1 ? SL_C1 : SL_C2) * 2 + p) * kPanelBytes, r, pc, cv);
            }
          }
          NVP_TL(issuer && it == 3, 16 * step + 8 * p + 1);   // TMEM loaded, activation math done
          if (step < 2 || TRAIN) store_rown<CW>(hbuf + p * kPanelBytes, r, pc, hv);
          if (step < 2) {
            store_rown<CW>(abuf + p * kPanelBytes, r, pc, av);
          } else {
#pragma unroll
            for (int i = 0; i < CW; ++i) {
              rgb0 = fmaf(av[i], s_wl[col + i], rgb0);
              rgb1 = fmaf(av[i], s_wl[H + col + i], rgb1);
              rgb2 = fmaf(av[i], s_wl[2 * H + col + i], rgb2);
            }
            if (p == 1 && sub > 0) { float* x = s_rgbx + ((sub - 1) * H + r) * 3; x[0] = rgb0; x[1] = rgb1; x[2] = rgb2; }
          }
          NVP_TL(issuer && it == 3, 16 * step + 8 * p + 2);
          fence_proxy_async_smem();   // st.shared operand / staging tiles -> visible to UMMA and TMA (async proxy)
          tcgen05_fence_before();
          asm volatile("bar.sync 2, %0;" ::"n"(EW * 32) : "memory");
          NVP_TL(issuer && it == 3, 16 * step + 8 * p + 3);
          if (issuer) {
            mbar_arrive(&panel_done[step * 2 + p]);   // first: the MMA warp is on the critical path, the stash stores are not
            if (TRAIN) {
              auto put = [&](int slot, const uint8_t* src) {
                bulk_s2g(st_base + (static_cast<size_t>(slot) * 2 + p) * kPanelBytes, src + p * kPanelBytes, kPanelBytes);
              };
              if (step == 0) { put(SL_H0, hbuf); put(SL_A0, abuf); }
              else if (step == 1) { put(SL_H1, hbuf); put(SL_A1, abuf); if (STAGE_SC) { put(SL_S1, sbuf); put(SL_C1, cbuf); } }
              else { put(SL_H2, hbuf); if (STAGE_SC) { put(SL_S2, sbuf); put(SL_C2, cbuf); } }
              bulk_commit();
            }
          }
        }
        if (step == 2 && sub == 0 && valid) {
#pragma unroll
          for (int q = 0; q < EW / 4 - 1; ++q) {   // partial sums of the other column slices
            rgb0 += s_rgbx[(q * H + r) * 3]; rgb1 += s_rgbx[(q * H + r) * 3 + 1]; rgb2 += s_rgbx[(q * H + r) * 3 + 2];
          }
          a.rgb[s * 3] = rgb0 + s_bl[0];
          a.rgb[s * 3 + 1] = rgb1 + s_bl[1];
          a.rgb[s * 3 + 2] = rgb2 + s_bl[2];
        }
      }
    }
    if (TRAIN && issuer) bulk_wait_all0();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ==========================================================================================
// Backward (dgrad) kernel
// ==========================================================================================
// Per 128-sample tile, with every gradient pre-multiplied by the power-of-two loss scale gs so that fp16
// operands stay in range:
//   P2  drgb = gs * dL/drgb ; da2 = drgb Wl ; dsp2 = da2*h2*cos2 ; dm2 = da2*sin2*lrelu'(h2)
//   S2  da1 = dsp2 Ws2 ; dh1 = dm2 W2h ; dz  = dm2 W2z         (tcgen05, accumulators in TMEM)
//   P1  dsp1 = da1*h1*cos1 ; dm1 = (dh1 + da1*sin1)*lrelu'(h1)
//   S1  da0 = dsp1 Ws1 ; dh0 = dm1 W1h ; dz += dm1 W1z
//   P0  dsp0 = da0*h0*cos0*w0 ; dm0 = (dh0 + da0*sin0)*lrelu'(h0)
//   S0  dz += dm0 W0z
//   PZ  dz -> HBM (fp16 MMA tile format, still scaled by gs; the grid scatter un-scales)
// Data movement is all TMA: each P-phase works on one 64-column panel whose stashed activations
// (h, sin, cos) were bulk-loaded into a shared-memory staging set, computes IN PLACE (h -> dm,
// cos -> dsp, sin -> a2), after which the same buffers are the K-major A operands of the dgrad GEMMs,
// the MN-major A operands of the small reductions, and the source of the bulk stores that hand
// dm/dsp to the wgrad kernel.  Two staging sets alternate between the two panels of a step.
// The small reductions over samples (dWl, db_siren, dw/db of SIREN layer 0) are "skinny" MN-major MMAs
//   acc[128 features, 16] += X^T R   against a per-row panel R = [drgb | 1 | 1 | 1,tau]
// whose 16-column blocks select the output column, accumulated across tiles in one persistent TMEM
// accumulator and flushed once per CTA.
enum { DP_M0 = 0, DP_M1, DP_M2, DP_S1, DP_S2, DP_COUNT };
constexpr int kBwdPanels = 14;
enum { ROLE_H = 0, ROLE_S = 1, ROLE_C = 2 };
constexpr uint32_t kSetBytes = 3 * kPanelBytes;

struct BwdArgs {
  const uint8_t* wpk;
  uint32_t poff[kBwdPanels], pbytes[kBwdPanels];
  const uint8_t* stash;
  const float* tau;
  const float* rgb;
  const uint8_t* gt;
  const float* dout;
  const float* gscale;   // device: [0] gs, [1] 1/gs, [2] gs*2/(3*n_global)
  const float* siren_w0;
  const float* siren_b0;
  const float* last_w;
  float w0;
  uint8_t* dpre;         // [tile][DP_COUNT][2 panels]
  uint8_t* dz16t;        // [tile][ZP/64 panels] fp16
  float* loss_sum;
  float* g_last_w; float* g_last_b; float* g_siren_b1; float* g_siren_b2; float* g_siren_w0; float* g_siren_b0;
  int64_t n;
  int n_tiles, ZP, NZ, nstage;   // NZ = dz GEMM width = round_up(Z+1, 16) <= ZP
  int stage_dz;                  // 1: dz tiles staged in smem and bulk-stored; 0: per-lane stores (no smem left)
  uint32_t stage_bytes;
};

struct BwdSmem { uint32_t sets, rp, dzst, ring, consts, bars, total; };
__host__ __device__ inline BwdSmem bwd_smem_layout(int nstage, uint32_t stage_bytes, int ZP, int stage_dz) {
  BwdSmem s;
  uint32_t o = 0;
  s.sets = o; o += 2 * kSetBytes;
  s.rp = o; o += kPanelBytes;
  s.dzst = o; if (stage_dz) o += (ZP / 64) * kPanelBytes;
  s.ring = o; o += nstage * stage_bytes;
  s.consts = o; o += 5 * H * 4;   // ws0[H] bs0[H] (both pre-multiplied by w0) wl[3][H]
  s.bars = o; o += 64 * 8;
  s.total = o;
  return s;
}

__device__ __forceinline__ void load_row32(const uint8_t* panel, int r, int c0, float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 q = *reinterpret_cast<const uint4*>(panel + panel_chunk_offset(r, (c0 >> 3) + j));
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(h[k]);
      v[8 * j + 2 * k] = f.x;
      v[8 * j + 2 * k + 1] = f.y;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) mlp_backward_kernel(const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const BwdSmem L = bwd_smem_layout(a.nstage, a.stage_bytes, a.ZP, a.stage_dz);
  uint8_t* sets = smem + L.sets;
  uint8_t* rp = smem + L.rp;
  uint8_t* dzst = smem + L.dzst;
  uint8_t* ring = smem + L.ring;
  float* s_ws0 = reinterpret_cast<float*>(smem + L.consts);
  float* s_bs0 = s_ws0 + H;
  float* s_wl = s_ws0 + 2 * H;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* wfull = bars;
  uint64_t* wempty = bars + 16;
  uint64_t* ld_full = bars + 32;      // [2] stash panels landed in set p
  uint64_t* panel_done = bars + 34;   // [2] epilogue finished panel p of the current step
  uint64_t* kp_done = bars + 36;      // [2] MMAs reading set p have completed
  uint64_t* acc_full = bars + 38;
  __shared__ uint32_t s_tmem;
  auto buf = [&](int set, int role) { return sets + set * kSetBytes + role * kPanelBytes; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < H; i += kThreads) { s_ws0[i] = __ldg(a.siren_w0 + i) * a.w0; s_bs0[i] = __ldg(a.siren_b0 + i) * a.w0; }
  for (int i = tid; i < 3 * H; i += kThreads) s_wl[i] = __ldg(a.last_w + i);
  if (tid == 0) {
    for (int i = 0; i < a.nstage; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&ld_full[i], 1); mbar_init(&panel_done[i], 1); mbar_init(&kp_done[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t acc_da = tmem, acc_dh = tmem + 128, acc_dz = tmem + 256, acc_sk = tmem + 256 + a.NZ;

  if (warp == 0) {
    // ================= TMA producer: weight ring =================
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        for (int i = 0; i < kBwdPanels; ++i, ++g) {
          const uint32_t st = g % a.nstage, ph = (g / a.nstage) & 1;
          mbar_wait(&wempty[st], ph ^ 1);
          mbar_arrive_expect_tx(&wfull[st], a.pbytes[i]);
          bulk_g2s(ring + st * a.stage_bytes, a.wpk + a.poff[i], a.pbytes[i], &wfull[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc_h = umma_idesc_f16(kTile, H, false, false);
      const uint32_t idesc_z = umma_idesc_f16(kTile, a.NZ, false, false);
      const uint32_t idesc_sk = umma_idesc_f16(H, 16, true, true);
      uint32_t g = 0, n_step = 0;
      bool sk_started = false;
      auto gemm = [&](uint8_t* a_panel, uint32_t acc, uint32_t idesc, bool accumulate) {
        const uint32_t st = g % a.nstage, ph = (g / a.nstage) & 1;
        mbar_wait(&wfull[st], ph);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(a_panel), b_addr = smem_u32(ring + st * a.stage_bytes);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          umma_f16_ss(acc, umma_desc_kmajor(a_addr, kk), umma_desc_kmajor(b_addr, kk), idesc, accumulate ? 1u : 0u);
          accumulate = true;
        }
        umma_commit(&wempty[st]);
        ++g;
      };
      auto skinny = [&](int role, int block) {   // A = role buffers of set 0 (features 0-63) and set 1 (64-127)
        const uint32_t a_addr = smem_u32(buf(0, role)), b_addr = smem_u32(rp) + static_cast<uint32_t>(block) * 32u;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          umma_f16_ss(acc_sk, umma_desc_mnmajor(a_addr, kk, kSetBytes), umma_desc_mnmajor(b_addr, kk, kPanelBytes), idesc_sk,
                      sk_started ? 1u : 0u);
          sk_started = true;
        }
      };
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        for (int step = 2; step >= 0; --step, ++n_step) {
          const uint32_t par = n_step & 1;
          // K panel 0 of dz may start as soon as panel 0 of dm is ready.  The da/dh GEMMs overwrite accumulators
          // whose upper 64 columns the epilogue is still reading during its panel-1 phase (no TMEM left to
          // double-buffer them), so they wait for panel 1.
          mbar_wait(&panel_done[0], par);
          tcgen05_fence_after();
          gemm(buf(0, ROLE_H), acc_dz, idesc_z, step != 2);
          mbar_wait(&panel_done[1], par);
          tcgen05_fence_after();
          if (step == 2) { skinny(ROLE_S, 0); skinny(ROLE_C, 1); }
          else if (step == 1) skinny(ROLE_C, 2);
          else skinny(ROLE_C, 3);
          if (step > 0) {
            gemm(buf(0, ROLE_C), acc_da, idesc_h, false);
            gemm(buf(0, ROLE_H), acc_dh, idesc_h, false);
          }
          umma_commit(&kp_done[0]);
          if (step > 0) {
            gemm(buf(1, ROLE_C), acc_da, idesc_h, true);
            gemm(buf(1, ROLE_H), acc_dh, idesc_h, true);
          }
          gemm(buf(1, ROLE_H), acc_dz, idesc_z, true);
          umma_commit(&kp_done[1]);
          umma_commit(acc_full);
        }
      }
    }
  } else {
    // ================= epilogue warps =================
    const int quarter = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const int pc = sub * 32;               // my 32 columns inside the current 64-column panel
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const bool issuer = (tid == 64);
    const float gs = __ldg(a.gscale), inv_gs = __ldg(a.gscale + 1), loss_mult = __ldg(a.gscale + 2);
    float loss_acc = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
    uint32_t n_phase[2] = {0, 0};   // completed uses of set 0 / set 1 (parity of ld_full / kp_done)
    uint32_t n_acc = 0;
    bool any = false;

    // stash panels needed by phase (step, p), bulk-loaded into staging set p
    auto issue_loads = [&](int tile, int step, int p) {
      const uint8_t* st_base = a.stash + (static_cast<size_t>(tile) * SL_COUNT) * 2 * kPanelBytes;
      auto src = [&](int slot) { return st_base + (static_cast<size_t>(slot) * 2 + p) * kPanelBytes; };
      if (step == 2) {
        mbar_arrive_expect_tx(&ld_full[p], 3 * kPanelBytes);
        bulk_g2s(buf(p, ROLE_H), src(SL_H2), kPanelBytes, &ld_full[p]);
        bulk_g2s(buf(p, ROLE_S), src(SL_S2), kPanelBytes, &ld_full[p]);
        bulk_g2s(buf(p, ROLE_C), src(SL_C2), kPanelBytes, &ld_full[p]);
      } else if (step == 1) {
        mbar_arrive_expect_tx(&ld_full[p], 3 * kPanelBytes);
        bulk_g2s(buf(p, ROLE_H), src(SL_H1), kPanelBytes, &ld_full[p]);
        bulk_g2s(buf(p, ROLE_S), src(SL_S1), kPanelBytes, &ld_full[p]);
        bulk_g2s(buf(p, ROLE_C), src(SL_C1), kPanelBytes, &ld_full[p]);
      } else {
        mbar_arrive_expect_tx(&ld_full[p], kPanelBytes);
        bulk_g2s(buf(p, ROLE_H), src(SL_H0), kPanelBytes, &ld_full[p]);
      }
    };
    const int first_tile = blockIdx.x;
    if (issuer && first_tile < a.n_tiles) { issue_loads(first_tile, 2, 0); issue_loads(first_tile, 2, 1); }

    for (int tile = first_tile; tile < a.n_tiles; tile += gridDim.x) {
      any = true;
      const int next_tile = tile + gridDim.x;
      const int64_t s = static_cast<int64_t>(tile) * kTile + r;
      const bool valid = s < a.n;
      const float tau = valid ? __ldg(a.tau + s) : 0.0f;
      uint8_t* dp_base = a.dpre + (static_cast<size_t>(tile) * DP_COUNT) * 2 * kPanelBytes;

      // drgb of my row (scaled by gs) and the per-row panel R of the small reductions
      float d0 = 0.f, d1 = 0.f, d2 = 0.f;
      if (valid) {
        if (a.dout != nullptr) {
          d0 = __ldg(a.dout + s * 3) * gs; d1 = __ldg(a.dout + s * 3 + 1) * gs; d2 = __ldg(a.dout + s * 3 + 2) * gs;
        } else {
          const float e0 = a.rgb[s * 3] - (static_cast<float>(a.gt[s * 3]) - 127.5f) / 127.5f;
          const float e1 = a.rgb[s * 3 + 1] - (static_cast<float>(a.gt[s * 3 + 1]) - 127.5f) / 127.5f;
          const float e2 = a.rgb[s * 3 + 2] - (static_cast<float>(a.gt[s * 3 + 2]) - 127.5f) / 127.5f;
          if (sub == 0) loss_acc += e0 * e0 + e1 * e1 + e2 * e2;
          d0 = e0 * loss_mult; d1 = e1 * loss_mult; d2 = e2 * loss_mult;
        }
      }
      if (sub == 0) {
        b0 += d0; b1 += d1; b2 += d2;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(rp + panel_chunk_offset(r, 0)) = make_uint4(pack_half2(d0, d1), pack_half2(d2, 0.f), 0u, 0u);
        *reinterpret_cast<uint4*>(rp + panel_chunk_offset(r, 1)) = zero;
        *reinterpret_cast<uint4*>(rp + panel_chunk_offset(r, 2)) = make_uint4(0u, pack_half2(0.f, 1.f), 0u, 0u);
        *reinterpret_cast<uint4*>(rp + panel_chunk_offset(r, 3)) = zero;
        *reinterpret_cast<uint4*>(rp + panel_chunk_offset(r, 4)) = make_uint4(0u, 0u, pack_half2(1.f, 0.f), 0u);
        *reinterpret_cast<uint4*>(rp + panel_chunk_offset(r, 5)) = zero;
        *reinterpret_cast<uint4*>(rp + panel_chunk_offset(r, 6)) = make_uint4(0u, 0u, pack_half2(0.f, 1.f), pack_half2(tau, 0.f));
        *reinterpret_cast<uint4*>(rp + panel_chunk_offset(r, 7)) = zero;
      }

#pragma unroll 1
      for (int step = 2; step >= 0; --step) {
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
          const int col = p * 64 + pc;
          uint8_t* bh = buf(p, ROLE_H);
          uint8_t* bs = buf(p, ROLE_S);
          uint8_t* bc = buf(p, ROLE_C);
          if (p == 0 && step < 2) {
            // da/dh of this layer are complete; the MMAs that read staging set 1 are done as well
            mbar_wait(acc_full, n_acc & 1); ++n_acc;
            tcgen05_fence_after();
            if (issuer) {
              bulk_wait_read0();                       // set 1's dm/dsp bulk stores have left shared memory
              issue_loads(tile, step, 1);
            }
          }
          mbar_wait(&ld_full[p], n_phase[p] & 1);
          float hv[32], sv[32], cv[32], o[32];
          load_row32(bh, r, pc, hv);
          if (step == 2) {
            load_row32(bs, r, pc, sv);
            load_row32(bc, r, pc, cv);
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = sv[i] * hv[i];     // a2, operand of the dWl reduction
            store_row32(bs, r, pc, o);
            float da[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) da[i] = d0 * s_wl[col + i] + d1 * s_wl[H + col + i] + d2 * s_wl[2 * H + col + i];
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = da[i] * hv[i] * cv[i];
            store_row32(bc, r, pc, o);
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = da[i] * sv[i] * (hv[i] > 0.f ? 1.f : 0.01f);
            store_row32(bh, r, pc, o);
          } else {
            uint32_t va[32], vh[32];
            tmem_ld32(acc_da + lane_base + col, va);
            tmem_ld32(acc_dh + lane_base + col, vh);
            float post = 1.0f;
            if (step == 1) {
              load_row32(bs, r, pc, sv);
              load_row32(bc, r, pc, cv);
            } else {
              post = a.w0;
#pragma unroll
              for (int i = 0; i < 32; ++i) fast_sincos(fmaf(tau, s_ws0[col + i], s_bs0[col + i]), sv[i], cv[i]);
            }
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(va[i]) * hv[i] * cv[i] * post;
            store_row32(bc, r, pc, o);
#pragma unroll
            for (int i = 0; i < 32; ++i)
              o[i] = (__uint_as_float(vh[i]) + __uint_as_float(va[i]) * sv[i]) * (hv[i] > 0.f ? 1.f : 0.01f);
            store_row32(bh, r, pc, o);
          }
          ++n_phase[p];
          fence_proxy_async_smem();
          tcgen05_fence_before();
          asm volatile("bar.sync 2, %0;" ::"n"(kEpiWarps * 32) : "memory");
          if (issuer) {
            mbar_arrive(&panel_done[p]);   // first: the MMA warp is on the critical path, the dpre stores are not
            auto put = [&](int slot, const uint8_t* src) {
              bulk_s2g(dp_base + (static_cast<size_t>(slot) * 2 + p) * kPanelBytes, src, kPanelBytes);
            };
            if (step == 2) { put(DP_S2, bc); put(DP_M2, bh); }
            else if (step == 1) { put(DP_S1, bc); put(DP_M1, bh); }
            else put(DP_M0, bh);
            bulk_commit();
            if (p == 1) {
              // staging set 0 is free once the small reductions (the last MMAs reading it) have completed and its
              // bulk stores have been read: refill it for the next phase that uses it.
              const bool more = step > 0 || next_tile < a.n_tiles;
              if (more) {
                mbar_wait(&kp_done[0], (n_phase[0] - 1) & 1);
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                if (step > 0) issue_loads(tile, step - 1, 0); else issue_loads(next_tile, 2, 0);
              }
            }
          }
        }
      }

      // ---------------- PZ ----------------
      mbar_wait(acc_full, n_acc & 1); ++n_acc;
      tcgen05_fence_after();
      if (issuer && next_tile < a.n_tiles) {
        bulk_wait_read0();
        issue_loads(next_tile, 2, 1);
      }
      const int zp_panels = a.ZP >> 6;
      uint8_t* dz_tile = a.dz16t + static_cast<size_t>(tile) * zp_panels * kPanelBytes;
      for (int q = 0; q < zp_panels; ++q) {
        if (q * 64 + pc >= a.NZ) break;        // columns >= NZ hold no gradient (never read by the scatter)
        uint32_t v[32];
        float f[32];
        tmem_ld32(acc_dz + lane_base + q * 64 + pc, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        store_row32((a.stage_dz ? dzst : dz_tile) + q * kPanelBytes, r, pc, f);
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      asm volatile("bar.sync 2, %0;" ::"n"(kEpiWarps * 32) : "memory");
      if (issuer && a.stage_dz) {
        bulk_s2g(dz_tile, dzst, zp_panels * kPanelBytes);
        bulk_commit();
      }
    }
    if (issuer) bulk_wait_all0();
    // ---------------- flush of the per-CTA reductions ----------------
    if (any && sub == 0) {
      tcgen05_fence_after();
      uint32_t v[16];
      tmem_ld16(acc_sk + lane_base, v);
      tmem_ld_wait();
      const int j = r;  // feature index
      if (a.g_last_w) {
        atomicAdd(a.g_last_w + j, __uint_as_float(v[0]) * inv_gs);
        atomicAdd(a.g_last_w + H + j, __uint_as_float(v[1]) * inv_gs);
        atomicAdd(a.g_last_w + 2 * H + j, __uint_as_float(v[2]) * inv_gs);
      }
      if (a.g_siren_b2) atomicAdd(a.g_siren_b2 + j, __uint_as_float(v[3]) * inv_gs);
      if (a.g_siren_b1) atomicAdd(a.g_siren_b1 + j, __uint_as_float(v[4]) * inv_gs);
      if (a.g_siren_b0) atomicAdd(a.g_siren_b0 + j, __uint_as_float(v[5]) * inv_gs);
      if (a.g_siren_w0) atomicAdd(a.g_siren_w0 + j, __uint_as_float(v[6]) * inv_gs);
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) {
        loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o2);
        b0 += __shfl_xor_sync(0xffffffffu, b0, o2);
        b1 += __shfl_xor_sync(0xffffffffu, b1, o2);
        b2 += __shfl_xor_sync(0xffffffffu, b2, o2);
      }
      if (lane == 0) {
        if (a.loss_sum && a.dout == nullptr) atomicAdd(a.loss_sum, loss_acc);
        if (a.g_last_b) {
          atomicAdd(a.g_last_b, b0 * inv_gs); atomicAdd(a.g_last_b + 1, b1 * inv_gs); atomicAdd(a.g_last_b + 2, b2 * inv_gs);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ==========================================================================================
// Weight-gradient kernel: dW[out, in] = sum over samples of dpre[s, out] * act[s, in]
// ==========================================================================================
// Both operands are MN-major views of stored tiles (K = samples).  The accumulators of all seven
// products do not fit one SM's TMEM (512 columns), so CTAs come in two kinds, each persistent over
// all half-tiles (64 samples per pipeline stage) with its products resident in TMEM:
//   kind A: dW1h = dm1^T h0 | dW1z = dm1^T z | dWs1 = dsp1^T a0
//   kind B: dW0z = dm0^T z  | dW2h = dm2^T h1 | dW2z = dm2^T z | dWs2 = dsp2^T a1
// Column Z of every z product is the sum of dm over samples (z carries a constant-1 column) = the
// modulator bias gradient.  Accumulators are flushed once per CTA with fp32 atomics, un-scaled by 1/gs.
constexpr int kWgThreads = 192;
constexpr int kHalfPanelBytes = kPanelBytes / 2;  // 64 rows

// Operand sources: 0..DP_COUNT-1 = dpre slot, 16+slot = stash slot, 64 = latent z.
enum { WG_SRC_STASH = 16, WG_SRC_Z = 64 };
// Gradient targets of a product.
enum { WG_W0Z = 0, WG_W1H, WG_W1Z, WG_W2H, WG_W2Z, WG_WS1, WG_WS2 };

struct WgProduct { int a_hp, b_hp, is_z, tmem_col, target; };   // operand positions in half-panels inside a stage
struct WgKind {
  int n_ops, op_src[6], n_hp;   // operands loaded per stage and total half-panels
  int n_prod;
  WgProduct prod[4];
  int cta_begin, cta_end;
};
struct WgArgs {
  const uint8_t* dpre;
  const uint8_t* stash;
  const uint8_t* z16t;
  const float* gscale;
  float* g_mod_w[3];
  float* g_mod_b[3];
  float* g_siren_w[3];
  int n_units;   // half tiles
  int stash_slots;   // stash tiles per 128 samples (SL_COUNT, or FS_COUNT behind the fused kernel; slots 0-3 are h0 a0 h1 a1)
  int KZ, Z, ZP;
  int n_kinds;
  WgKind kind[3];
};

__global__ void __launch_bounds__(kWgThreads, 1) mlp_wgrad_kernel(const WgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[2], empty[2], done;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int ki = 0;
  while (ki + 1 < a.n_kinds && static_cast<int>(blockIdx.x) >= a.kind[ki].cta_end) ++ki;
  const WgKind& K = a.kind[ki];
  const int rank = blockIdx.x - K.cta_begin;
  const int stride = K.cta_end - K.cta_begin;
  const uint32_t stage_bytes = static_cast<uint32_t>(K.n_hp) * kHalfPanelBytes;

  if (tid == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&empty[0], 1); mbar_init(&empty[1], 1); mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const int my_units = rank < a.n_units ? (a.n_units - rank + stride - 1) / stride : 0;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int u = rank; u < a.n_units; u += stride, ++it) {
        const int st = it & 1, ph = (it >> 1) & 1;
        mbar_wait(&empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&full[st], stage_bytes);
        uint8_t* dst = smem + st * stage_bytes;
        const int tile = u >> 1;
        const uint32_t hoff = static_cast<uint32_t>(u & 1) * kHalfPanelBytes;  // rows 0-63 or 64-127 of each panel
        const uint8_t* dp = a.dpre + static_cast<size_t>(tile) * DP_COUNT * 2 * kPanelBytes;
        const uint8_t* sb = a.stash + static_cast<size_t>(tile) * a.stash_slots * 2 * kPanelBytes;
        const uint8_t* zt = a.z16t + static_cast<size_t>(tile) * a.KZ * kPanelBytes;
        for (int o = 0; o < K.n_ops; ++o) {
          const int src = K.op_src[o];
          const uint8_t* base = src == WG_SRC_Z ? zt : (src >= WG_SRC_STASH ? sb + static_cast<size_t>(src - WG_SRC_STASH) * 2 * kPanelBytes
                                                                           : dp + static_cast<size_t>(src) * 2 * kPanelBytes);
          const int np = src == WG_SRC_Z ? a.KZ : 2;
          for (int q = 0; q < np; ++q) {
            bulk_g2s(dst, base + static_cast<size_t>(q) * kPanelBytes + hoff, kHalfPanelBytes, &full[st]);
            dst += kHalfPanelBytes;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_h = umma_idesc_f16(H, H, true, true);
      const uint32_t idesc_z = umma_idesc_f16(H, a.ZP, true, true);
      int it = 0;
      for (int u = rank; u < a.n_units; u += stride, ++it) {
        const int st = it & 1, ph = (it >> 1) & 1;
        mbar_wait(&full[st], ph);
        tcgen05_fence_after();
        const uint32_t sb = smem_u32(smem + st * stage_bytes);
        for (int pi = 0; pi < K.n_prod; ++pi) {
          const WgProduct& P = K.prod[pi];
          const uint32_t a_addr = sb + static_cast<uint32_t>(P.a_hp) * kHalfPanelBytes;
          const uint32_t b_addr = sb + static_cast<uint32_t>(P.b_hp) * kHalfPanelBytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16_ss(tmem + P.tmem_col, umma_desc_mnmajor(a_addr, kk, kHalfPanelBytes),
                        umma_desc_mnmajor(b_addr, kk, kHalfPanelBytes), P.is_z ? idesc_z : idesc_h, (it > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&empty[st]);
      }
      umma_commit(&done);
    }
  } else if (my_units > 0) {
    // flush warps (4): thread = output feature row
    mbar_wait(&done, 0);
    tcgen05_fence_after();
    const int quarter = warp & 3;
    const int j = quarter * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const float inv_gs = __ldg(a.gscale + 1);
    for (int pi = 0; pi < K.n_prod; ++pi) {
      const WgProduct& P = K.prod[pi];
      float* gw; float* gb = nullptr; int ld, col_off = 0;
      switch (P.target) {
        case WG_W0Z: gw = a.g_mod_w[0]; gb = a.g_mod_b[0]; ld = a.Z; break;
        case WG_W1H: gw = a.g_mod_w[1]; ld = H + a.Z; break;
        case WG_W1Z: gw = a.g_mod_w[1]; gb = a.g_mod_b[1]; ld = H + a.Z; col_off = H; break;
        case WG_W2H: gw = a.g_mod_w[2]; ld = H + a.Z; break;
        case WG_W2Z: gw = a.g_mod_w[2]; gb = a.g_mod_b[2]; ld = H + a.Z; col_off = H; break;
        case WG_WS1: gw = a.g_siren_w[1]; ld = H; break;
        default:     gw = a.g_siren_w[2]; ld = H; break;
      }
      const int ncol = P.is_z ? a.Z + 1 : H;   // column Z of a z product = sum of dm over samples = bias gradient
      for (int c0 = 0; c0 < ncol; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + P.tmem_col + lane_base + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int c = c0 + i;
          const float val = __uint_as_float(v[i]) * inv_gs;
          if (!P.is_z || c < a.Z) { if (gw) atomicAdd(gw + static_cast<size_t>(j) * ld + col_off + c, val); }
          else if (c == a.Z) { if (gb) atomicAdd(gb + j, val); }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// Host-side plan: which products each CTA kind keeps resident in its 512 TMEM columns.
void build_wgrad_plan(WgArgs& wa, int sms) {
  const int ZP = wa.ZP, KZ = wa.KZ;
  auto kind = [&](std::initializer_list<int> srcs, std::initializer_list<WgProduct> prods) {
    WgKind k{};
    int hp = 0;
    for (int src : srcs) { k.op_src[k.n_ops++] = src; hp += (src == WG_SRC_Z ? KZ : 2); }
    k.n_hp = hp;
    int col = 0;
    for (WgProduct p : prods) { p.tmem_col = col; col += p.is_z ? ZP : H; k.prod[k.n_prod++] = p; }
    return k;
  };
  // operand half-panel positions follow the order of `srcs`: each dpre/stash operand takes 2, z takes KZ.
  const WgKind A = kind({DP_M1, DP_S1, WG_SRC_STASH + SL_H0, WG_SRC_STASH + SL_A0, WG_SRC_Z},
                        {{0, 4, 0, 0, WG_W1H}, {0, 8, 1, 0, WG_W1Z}, {2, 6, 0, 0, WG_WS1}});
  if (2 * ZP + 256 <= 512) {
    const WgKind B = kind({DP_M0, DP_M2, DP_S2, WG_SRC_STASH + SL_H1, WG_SRC_STASH + SL_A1, WG_SRC_Z},
                          {{0, 10, 1, 0, WG_W0Z}, {2, 6, 0, 0, WG_W2H}, {2, 10, 1, 0, WG_W2Z}, {4, 8, 0, 0, WG_WS2}});
    wa.n_kinds = 2; wa.kind[0] = A; wa.kind[1] = B;
  } else {
    const WgKind B = kind({DP_M2, DP_S2, WG_SRC_STASH + SL_H1, WG_SRC_STASH + SL_A1, WG_SRC_Z},
                          {{0, 4, 0, 0, WG_W2H}, {0, 8, 1, 0, WG_W2Z}, {2, 6, 0, 0, WG_WS2}});
    const WgKind C = kind({DP_M0, WG_SRC_Z}, {{0, 2, 1, 0, WG_W0Z}});
    wa.n_kinds = 3; wa.kind[0] = A; wa.kind[1] = B; wa.kind[2] = C;
  }
  // CTAs per kind proportional to the bytes a kind streams per unit (every kind sweeps all units)
  int total_hp = 0;
  for (int k = 0; k < wa.n_kinds; ++k) total_hp += wa.kind[k].n_hp;
  int begin = 0, left = sms;
  for (int k = 0; k < wa.n_kinds; ++k) {
    int c = (k == wa.n_kinds - 1) ? left : std::max(1, sms * wa.kind[k].n_hp / total_hp);
    c = std::max(1, std::min(c, wa.n_units));
    wa.kind[k].cta_begin = begin; wa.kind[k].cta_end = begin + c;
    begin += c; left -= c;
  }
}

// ------------------------------------------------------------------------------------------
// Loss-scale plumbing
// ------------------------------------------------------------------------------------------
__global__ void set_gscale_kernel(float* g, float gs, float mult) { g[0] = gs; g[1] = 1.0f / gs; g[2] = mult; }
__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, unsigned int* out) {
  float m = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
__global__ void gscale_from_absmax_kernel(const unsigned int* mx, float* g) {
  const float m = __uint_as_float(*mx);
  int e = 0;
  if (m > 0.f && isfinite(m)) frexpf(m, &e);   // m = f * 2^e, f in [0.5,1)
  const float gs = ldexpf(1.0f, 1 - e);          // max|dout| * gs in [1, 2)
  g[0] = gs; g[1] = 1.0f / gs; g[2] = 0.f;
}

// Backward weight stream per tile, in MMA consumption order: for layer 2 then 1:
// [W_z^T 0][Ws^T 0][W_h^T 0][Ws^T 1][W_h^T 1][W_z^T 1]; then [W0z^T 0][W0z^T 1].  A K panel holds 64 output features of a
// [N = input features, 128] matrix.
int pack_backward_panels(const nvp_desc* d, const nvp_params* p, PackArgs& a, uint32_t base_off, BwdArgs* b) {
  const Dims m = make_dims(d);
  uint32_t off = base_off;
  int k = 0;
  auto add = [&](const float* src, int ld, int r0, int rvalid, int rows, int q) {
    b->poff[k] = off - base_off; b->pbytes[k] = static_cast<uint32_t>(rows) * 128u; ++k;
    add_panel(a, src, ld, 1, r0, 64 * q, rvalid, 64, rows, off);
  };
  for (int i = 2; i >= 1; --i) {
    add(p->mod_w[i], H + m.Z, H, m.Z, m.ZP, 0);
    for (int q = 0; q < 2; ++q) {
      add(p->siren_w[i], H, 0, H, H, q);
      add(p->mod_w[i], H + m.Z, 0, H, H, q);
    }
    add(p->mod_w[i], H + m.Z, H, m.Z, m.ZP, 1);
  }
  for (int q = 0; q < 2; ++q) add(p->mod_w[0], m.Z, 0, m.Z, m.ZP, q);
  return 0;
}

#include "mlp_fused.cuh"

struct TcWorkspace {
  uint8_t* wpk_fused;
  uint8_t* fconsts;
  uint8_t* wpk_fwd;
  uint8_t* wpk_bwd;
  uint8_t* z16t;
  uint8_t* stash;
  uint8_t* dpre;
  uint8_t* dz16t;
  float* rgb;
  float* gscale;
  void* binws;     // bucket state of the binned grid path (NULL: direct grid kernels)
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
  w.wpk_bwd = take(static_cast<size_t>(8) * kPanelBytes + static_cast<size_t>(6) * m.ZP * 128);
  w.wpk_fused = take(static_cast<size_t>(kFPanelsPerTile) * kPanelBytes);
  w.fconsts = take(sizeof(FusedConsts));
  w.z16t = take(static_cast<size_t>(tiles) * m.KZ * kPanelBytes);
  w.rgb = reinterpret_cast<float*>(take(static_cast<size_t>(tiles) * kTile * 3 * sizeof(float)));
  {
    LevelTab tab;
    size_t bin_bytes = 0;
    if (build_level_table(d, &tab, nullptr) == 0) bin_bytes = grid_bin_workspace_bytes(d, tab, n);
    w.binws = bin_bytes ? take(bin_bytes) : nullptr;
  }
  if (what == 1) {
    w.stash = take(static_cast<size_t>(tiles) * SL_COUNT * 2 * kPanelBytes);
    w.dpre = take(static_cast<size_t>(tiles) * DP_COUNT * 2 * kPanelBytes);
    w.dz16t = take(static_cast<size_t>(tiles) * m.KZ * kPanelBytes);
    w.gscale = reinterpret_cast<float*>(take(64));
  }
  w.total = off + 1024;
  return w;
}

// SM count of the current device (cached per device index: a process may drive several, possibly different, GPUs).
int num_sms() {
  constexpr int kMaxDev = 64;
  static std::atomic<int> cache[kMaxDev];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < kMaxDev) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < kMaxDev) cache[dev].store(sms, std::memory_order_relaxed);
  return sms;
}

constexpr int kSmemBudget = 227 * 1024 - 2048;

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
  bool stage_sc = train;
  a.nstage = 0;
  for (int pass = 0; pass < 2 && a.nstage == 0; ++pass) {
    for (int ns = 12; ns >= 2; --ns)
      if (static_cast<int>(fwd_smem_layout(m.KZ, ns, stage_sc).total) <= kSmemBudget) { a.nstage = ns; break; }
    if (a.nstage == 0) stage_sc = false;
  }
  NVP_CHECK(a.nstage >= 2, "latent too wide for the shared-memory plan of the tensor-core path");
  const size_t smem = fwd_smem_layout(m.KZ, a.nstage, stage_sc).total + 1024;
  const int grid = std::min(a.n_tiles, num_sms());
  static const int epi_warps = [] { const char* v = getenv("NVP_FWD_EPI_WARPS"); return (v && atoi(v) == 16) ? 16 : 8; }();   // measured: no difference (the kernel is latency-chain bound)
  ScopedKernelTimer timer(K_MLP_FWD, st);
  auto launch = [&](auto kernel, int threads) -> int {
    NVP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kernel<<<grid, threads, smem, st>>>(a);
    return 0;
  };
  int rc;
  if (epi_warps == 16) {
    if (!train) rc = launch(mlp_forward_kernel<false, false, 16>, 576);
    else if (stage_sc) rc = launch(mlp_forward_kernel<true, true, 16>, 576);
    else rc = launch(mlp_forward_kernel<true, false, 16>, 576);
  } else {
    if (!train) rc = launch(mlp_forward_kernel<false, false, 8>, 320);
    else if (stage_sc) rc = launch(mlp_forward_kernel<true, true, 8>, 320);
    else rc = launch(mlp_forward_kernel<true, false, 8>, 320);
  }
  if (rc) return rc;
  NVP_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int tc_timeline_read(unsigned long long* out, int n) {
#ifdef NVP_TIMELINE
  NVP_CUDA(cudaMemcpyFromSymbol(out, g_timeline, sizeof(unsigned long long) * std::min(n, 128)));
  return 0;
#else
  (void)out; (void)n;
  NVP_CHECK(false, "library built without the in-kernel timeline (make TIMELINE=1)");
#endif
}

size_t tc_workspace_bytes(const nvp_desc* d, int64_t n, int what) { return carve_tc(d, std::max<int64_t>(n, 1), what, nullptr).total; }

int tc_forward(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords, const float* tsteps,
               int64_t n, float* out_rgb, void* ws, size_t ws_bytes, cudaStream_t st, bool temporal_interp) {
  NVP_CHECK(ws_bytes >= tc_workspace_bytes(d, n, 0), "workspace too small (see nvp_workspace_bytes)");
  void* base = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
  const TcWorkspace w = carve_tc(d, n, 0, base);
  const Dims m = make_dims(d);
  int rc;
  if ((rc = pack_forward_weights(d, p, w.wpk_fwd, nullptr, st))) return rc;
  if (w.binws) {
    if ((rc = launch_grid_bin(d, tab, coords, n, m.KZ, w.binws, st))) return rc;
    if ((rc = launch_grid_gather_binned(d, tab, p, coords, n, w.z16t, m.KZ, w.binws, st, temporal_interp))) return rc;
  } else if ((rc = launch_grid_gather(d, tab, p, coords, n, nullptr, 0, w.z16t, m.KZ, st, temporal_interp))) {
    return rc;
  }
  return launch_forward(d, p, w, tsteps, n, out_rgb, false, st);
}

int tc_fwd_bwd(const nvp_desc* d, const LevelTab& tab, const nvp_params* p, const float* coords, const float* tsteps,
               const uint8_t* gt_u8, const float* dout, int64_t n, int64_t n_global, const nvp_grads* g, float* loss_sum,
               float* out_rgb, void* ws, size_t ws_bytes, cudaStream_t st) {
  NVP_CHECK(ws_bytes >= tc_workspace_bytes(d, n, 1), "workspace too small (see nvp_workspace_bytes)");
  const Dims m = make_dims(d);
  const int NZ = round_up(m.Z + 1, 16);
  NVP_CHECK(m.ZP <= 256 && 256 + NZ + 16 <= 512, "tensor-core backward: latent wider than 239 columns is not built");
  void* base = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
  const TcWorkspace w = carve_tc(d, n, 1, base);
  const int n_tiles = static_cast<int>((n + kTile - 1) / kTile);
  int rc;

  // 1. loss scale
  if (dout == nullptr) {
    const double target = 1.5 * static_cast<double>(n_global);
    const float gs = ldexpf(1.0f, static_cast<int>(lrint(log2(target))));
    set_gscale_kernel<<<1, 1, 0, st>>>(w.gscale, gs, gs * 2.0f / (3.0f * static_cast<float>(n_global)));
    NVP_LAUNCH_CHECK();
  } else {
    unsigned int* mx = reinterpret_cast<unsigned int*>(w.gscale + 8);
    NVP_CUDA(cudaMemsetAsync(mx, 0, sizeof(unsigned int), st));
    absmax_kernel<<<std::min<int64_t>(1024, (3 * n + 255) / 256), 256, 0, st>>>(dout, 3 * n, mx);
    NVP_LAUNCH_CHECK();
    gscale_from_absmax_kernel<<<1, 1, 0, st>>>(mx, w.gscale);
    NVP_LAUNCH_CHECK();
  }

  // The fused forward+backward kernel covers latents of up to 127 columns (two 64-wide panels: config S); wider ones
  // (config L) take the three-kernel path.  NVP_MLP_FUSED=0 forces the latter (A/B measurements).
  static const bool fused_enabled = [] { const char* v = getenv("NVP_MLP_FUSED"); return !(v && atoi(v) == 0); }();
  const bool fused = fused_enabled && m.KZ == 2 && m.Z + 2 <= 128;   // two constant-1 latent columns must fit
  float* rgb = out_rgb ? out_rgb : w.rgb;
  BwdArgs b{};

  // 2. weights -> fp16 panels
  if (fused) { if ((rc = pack_fused_weights(d, p, w.wpk_fused, st))) return rc; }
  else if ((rc = pack_forward_weights(d, p, w.wpk_fwd, &b, st, w.wpk_bwd))) return rc;

  // 3. positional features (the gather zero-fills the rows past n of the last tile)
  if (w.binws) {
    if ((rc = launch_grid_bin(d, tab, coords, n, m.KZ, w.binws, st))) return rc;
    if ((rc = launch_grid_gather_binned(d, tab, p, coords, n, w.z16t, m.KZ, w.binws, st))) return rc;
  } else if ((rc = launch_grid_gather(d, tab, p, coords, n, nullptr, 0, w.z16t, m.KZ, st))) {
    return rc;
  }

  if (fused) {
    // 4+5. fused forward + loss + backward
    FusedArgs f{};
    f.wpk = w.wpk_fused; f.z16t = w.z16t; f.tau = tsteps; f.w0 = d->w0_first;
    {
      // epilogue constants -> a slot of constant memory (stream-ordered device-to-device copy)
      static std::atomic<unsigned> next_slot{0};
      f.cslot = static_cast<int>(next_slot.fetch_add(1) % kFConstSlots);
      FusedConsts* staged = reinterpret_cast<FusedConsts*>(w.fconsts);
      f.consts = reinterpret_cast<const float*>(staged);
      fused_consts_kernel<<<1, H, 0, st>>>(staged, p->siren_b[0], p->siren_b[1], p->siren_b[2], p->siren_w[0], p->last_w, p->last_b,
                                           d->w0_first);
      NVP_LAUNCH_CHECK();
#if NVP_FCONST == 1
      NVP_CUDA(cudaMemcpyToSymbolAsync(g_fused_consts, staged, sizeof(FusedConsts), static_cast<size_t>(f.cslot) * sizeof(FusedConsts),
                                       cudaMemcpyDeviceToDevice, st));
#endif
    }
    f.gt = gt_u8; f.dout = dout; f.gscale = w.gscale; f.rgb_out = out_rgb;
    f.stash = w.stash; f.dpre = w.dpre; f.dz16t = w.dz16t; f.loss_sum = loss_sum;
    f.g_last_w = g->last_w; f.g_last_b = g->last_b; f.g_siren_b1 = g->siren_b[1]; f.g_siren_b2 = g->siren_b[2];
    f.g_siren_w0 = g->siren_w[0]; f.g_siren_b0 = g->siren_b[0];
    f.n = n; f.n_tiles = n_tiles;
    const size_t smem = fused_smem_layout().total + 1024;
    static_assert(fused_smem_layout().total + 1024 <= 227 * 1024, "fused kernel: shared-memory plan exceeds 227 KiB");
    NVP_CUDA(cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    ScopedKernelTimer timer(K_MLP_FUSED, st);
    mlp_fused_kernel<<<std::min(n_tiles, num_sms()), kFThreads, smem, st>>>(f);
    NVP_LAUNCH_CHECK();
  } else {
    // 4. fused forward (keeps the activation stash)
    if ((rc = launch_forward(d, p, w, tsteps, n, rgb, true, st))) return rc;

    // 5. fused backward
    b.wpk = w.wpk_bwd; b.stash = w.stash; b.tau = tsteps; b.rgb = rgb; b.gt = gt_u8; b.dout = dout; b.gscale = w.gscale;
    b.siren_w0 = p->siren_w[0]; b.siren_b0 = p->siren_b[0]; b.last_w = p->last_w; b.w0 = d->w0_first;
    b.dpre = w.dpre; b.dz16t = w.dz16t; b.loss_sum = loss_sum;
    b.g_last_w = g->last_w; b.g_last_b = g->last_b; b.g_siren_b1 = g->siren_b[1]; b.g_siren_b2 = g->siren_b[2];
    b.g_siren_w0 = g->siren_w[0]; b.g_siren_b0 = g->siren_b[0];
    b.n = n; b.n_tiles = n_tiles; b.ZP = m.ZP; b.NZ = NZ;
    b.stage_bytes = static_cast<uint32_t>(m.ZP) * 128u;
    b.nstage = 0;
    for (b.stage_dz = 1; b.stage_dz >= 0 && b.nstage == 0; --b.stage_dz) {
      for (int ns = 12; ns >= 3; --ns)
        if (static_cast<int>(bwd_smem_layout(ns, b.stage_bytes, m.ZP, b.stage_dz).total) <= kSmemBudget) { b.nstage = ns; break; }
      if (b.nstage) break;
    }
    NVP_CHECK(b.nstage >= 2, "latent too wide for the shared-memory plan of the tensor-core backward");
    const size_t smem = bwd_smem_layout(b.nstage, b.stage_bytes, m.ZP, b.stage_dz).total + 1024;
    NVP_CUDA(cudaFuncSetAttribute(mlp_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    ScopedKernelTimer timer(K_MLP_BWD, st);
    mlp_backward_kernel<<<std::min(n_tiles, num_sms()), kThreads, smem, st>>>(b);
    NVP_LAUNCH_CHECK();
  }

  // 6. scatter-add into the grids (dz is still multiplied by gs).  It runs before the weight-gradient kernel so that a
  //    multi-GPU host can start the collective on the grid gradients (the bulk of the all-reduce) while wgrad computes.
  if (w.binws) rc = launch_grid_scatter_binned(d, tab, coords, n, w.dz16t, m.KZ, 1.0f, w.gscale + 1, g, w.binws, st);
  else rc = launch_grid_scatter(d, tab, coords, n, nullptr, 0, w.dz16t, m.KZ, 1.0f, w.gscale + 1, g, st);
  if (rc) return rc;
  // With an event registered the host is about to run a collective next to the weight-gradient kernel: that kernel is
  // persistent with ~200 KB of shared memory per CTA, so a few SMs are left free for the collective's CTAs.
  int comm_sms = 0;
  if (cudaEvent_t ev = take_grid_event()) {
    NVP_CUDA(cudaEventRecord(ev, st));
    const char* v = getenv("NVP_COMM_SMS");
    comm_sms = std::max(0, std::min(num_sms() / 2, v && *v ? atoi(v) : 16));
  }

  // 7. weight gradients
  {
    WgArgs wa{};
    wa.dpre = w.dpre; wa.stash = w.stash; wa.z16t = w.z16t; wa.gscale = w.gscale;
    for (int i = 0; i < 3; ++i) { wa.g_mod_w[i] = g->mod_w[i]; wa.g_mod_b[i] = g->mod_b[i]; wa.g_siren_w[i] = g->siren_w[i]; }
    wa.n_units = 2 * n_tiles; wa.stash_slots = fused ? static_cast<int>(FS_COUNT) : static_cast<int>(SL_COUNT); wa.KZ = m.KZ; wa.Z = m.Z; wa.ZP = m.ZP;
    build_wgrad_plan(wa, num_sms() - comm_sms);
    int max_hp = 0;
    for (int k = 0; k < wa.n_kinds; ++k) max_hp = std::max(max_hp, wa.kind[k].n_hp);
    const size_t smem = 2 * static_cast<size_t>(max_hp) * kHalfPanelBytes + 1024;
    NVP_CHECK(static_cast<int>(smem) <= kSmemBudget + 1024 && m.ZP + 256 <= 512, "latent too wide for the wgrad plan");
    NVP_CUDA(cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    ScopedKernelTimer timer(K_MLP_WGRAD, st);
    mlp_wgrad_kernel<<<wa.kind[wa.n_kinds - 1].cta_end, kWgThreads, smem, st>>>(wa);
    NVP_LAUNCH_CHECK();
  }

  return 0;
}

}  // namespace nvp
