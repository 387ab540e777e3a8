// Fused forward + loss + backward (dgrad) kernel of the tensor-core path: one launch per training step replaces
// mlp_forward<train> + mlp_backward.  Included by mlp_tc.cu (shares its helpers and the operand tile format).
//
// Reference semantics: modulation.py:83-92,112-121 (forward), loss_functions.py:3 + training.py:47-48 (loss),
// training.py:74 (autograd backward).  Arithmetic is the tensor-core mode's: fp16 GEMM operands, fp32 accumulation
// in TMEM, fp32 epilogues, gradients carried with the power-of-two loss scale gs.
//
// Why one kernel: a 128-sample tile's activations never leave the SM between its forward and its backward.  What
// the backward needs is kept on chip -- h0 / h1 as fp16 tiles in shared memory (they are the next layer's A operand
// anyway), the SIREN pre-activation of layer 1 in TMEM, layer 2's sin / cos in registers (its backward is the
// second half of its forward epilogue) -- and only what the weight-gradient kernel consumes is written to HBM:
// h0, a0, h1, a1 and the five pre-activation gradients (9 tiles per 128 samples instead of 9 + 5 written and 9 read
// back).  There is no activation stash to re-load and no second weight-ring warm-up per tile.
//
// One persistent CTA per SM, one tile in flight, 20 warps (640 threads, 96 registers):
//   warp 0       TMA producer: weight panel ring (28 panels per tile, in MMA consumption order) + latent tile
//   warp 1       MMA issuer (one thread): tcgen05.mma M=128, N=128, K=16, accumulators in four 128-column TMEM slots
//   warps 2-3    reducers: column sums over the tile's samples (dW_last, SIREN bias gradients, layer-0 w/b gradients)
//                read from the fp16 operand tiles in shared memory, accumulated in fp32 registers across tiles
//   warps 4-19   epilogue: tcgen05.ld -> fp32 math -> fp16 operand tiles written in the UMMA layout (4 warps per TMEM
//                lane quarter, 16 columns of a 64-column panel per thread)
// Phases of a tile (epilogue) and the MMA groups between them:
//   P0 E(F0) h0,a0 | P1 E(F1) h1,a1 | P2 E(F2) h2,sin2,cos2,a2 -> rgb, loss, drgb | P3 dsp2,dm2 |
//   P4 E(B1) dsp1,dm1 | P5 E(B0) dsp0,dm0 | P6 dz -> HBM
//   F0 (issued one tile ahead, during P5 of the previous tile), F1, F2, G2 (da1,dh1,dz), G1 (da0,dh0,dz+=), G0 (dz+=)
// Every phase works panel by panel (64 columns) so the next MMA group starts on K panel 0 while panel 1 is computed.
//
// TMEM slots (128 fp32 columns each):  S0: m1, da1, da0   S1: m0, sp1 (kept from F1 until P4)   S2: m2, dh1, dh0
//                                      S3: sp2, dz
// Shared-memory tiles (32 KiB each):   Z: z -> a2 -> dsp1 -> dz staging     H0: h0 -> dm0     A0: a0 -> u/dsp2 -> dsp0
//                                      H1: h1 -> dm1      A1: a1 -> v/dm2 -> z of the NEXT tile (Z and A1 swap roles)
#pragma once

constexpr int kFEpiWarps = 16;
constexpr int kFEpiThreads = kFEpiWarps * 32;
constexpr int kFThreads = 128 + kFEpiThreads;   // warp 0 TMA, warp 1 MMA, warps 2-3 reducers, warps 4-19 epilogue
constexpr int kFCols = 256 / kFEpiWarps;        // columns of a 64-column panel one epilogue thread owns (16)
constexpr int kFSub = 64 / kFCols;              // epilogue warps per TMEM lane quarter (4)
constexpr int kFStages = 3;
constexpr int kFPanelsPerTile = 28;
constexpr uint32_t kFBuf = 2u * kPanelBytes;   // one [128 x 128] fp16 operand tile
enum { FS_H0 = 0, FS_A0, FS_H1, FS_A1, FS_COUNT };   // stash slots the weight-gradient kernel reads

// barrier slots
enum {
  FB_WFULL = 0, FB_WEMPTY = 3, FB_ZFULL = 6,
  FB_AF_P0 = 7, FB_AF_P1, FB_AF_P2, FB_AF_P4, FB_AF_P5, FB_AF_P6,
  FB_PD_P0 = 13, FB_PD_P1 = 15, FB_PD_P3 = 17, FB_PD_P4 = 19, FB_PD_P5 = 21, FB_PD_P6 = 23,   // panel done (one arrival per epilogue warp)
  FB_TF_P0 = 24, FB_TF_P4, FB_G2_DONE, FB_A2_READY, FB_RD_A2, FB_RD_DSP2, FB_RD_DSP1, FB_RD_DSP0,
  FB_SR_P0, FB_SR_P1, FB_SR_P3, FB_SR_P4, FB_SR_P5, FB_SR_P6,   // the bulk stores of that phase's tiles have left shared memory
  FB_COUNT
};

struct FusedArgs {
  const uint8_t* wpk;      // 28 weight panels in stream order (see pack_fused_weights)
  const uint8_t* z16t;     // latent tiles [tile][2 panels]
  const float* tau;
  int cslot;               // slot of g_fused_consts holding this call's epilogue constants (NVP_FCONST == 1)
  const float* consts;     // the same FusedConsts staged in global memory
  float w0;
  const uint8_t* gt;       // loss mode (dout == NULL)
  const float* dout;       // explicit upstream gradient (nvp_backward)
  const float* gscale;     // device: [0] gs, [1] 1/gs, [2] gs*2/(3*n_global)
  float* rgb_out;          // optional [n,3]
  uint8_t* stash;          // [tile][FS_COUNT][2 panels]
  uint8_t* dpre;           // [tile][DP_COUNT][2 panels]
  uint8_t* dz16t;          // [tile][2 panels]
  float* loss_sum;
  float* g_last_w; float* g_last_b; float* g_siren_b1; float* g_siren_b2; float* g_siren_w0; float* g_siren_b0;
  int64_t n;
  int n_tiles;
};

// Per-column constants of the epilogues, staged once per call by fused_consts_kernel and copied to shared memory by
// every CTA (or read from constant memory, see NVP_FCONST below - measured slower).
// The modulator biases are not here: they ride in the GEMMs (see pack_fused_weights).
struct FusedConsts {
  float bs[3][H];   // SIREN biases, layer 0 pre-multiplied by w0
  float ws0[H];     // net.layers.0.weight, pre-multiplied by w0
  float wl[3][H];   // net.last_layer.weight
  float bl[4];      // net.last_layer.bias
};
constexpr int kFConstSlots = 8;   // rotating slots: up to 8 fused calls of one process may be in flight on a device
__constant__ FusedConsts g_fused_consts[kFConstSlots];

__global__ void fused_consts_kernel(FusedConsts* out, const float* b0, const float* b1, const float* b2, const float* w_first,
                                    const float* last_w, const float* last_b, float w0) {
  const int i = threadIdx.x;
  if (i < H) {
    out->bs[0][i] = __ldg(b0 + i) * w0; out->bs[1][i] = __ldg(b1 + i); out->bs[2][i] = __ldg(b2 + i);
    out->ws0[i] = __ldg(w_first + i) * w0;
    out->wl[0][i] = __ldg(last_w + i); out->wl[1][i] = __ldg(last_w + H + i); out->wl[2][i] = __ldg(last_w + 2 * H + i);
  }
  if (i < 4) out->bl[i] = i < 3 ? __ldg(last_b + i) : 0.0f;
}

// NVP_FCONST: where the epilogue reads its constants from.  0 = a copy in shared memory (LDS.128, address uniform over the
// warp), 1 = constant memory (LDC).  Measured on B200: LDC with a register index is slower than the LDS it replaces
// (P3 of the timeline: 2.5 K -> 3.5 K cycles), so shared memory is the default.
#ifndef NVP_FCONST
#define NVP_FCONST 0
#endif
// NVP_ABL: timing-only ablations for scripts/fused_ablate.sh (results are wrong when set):
//   1 constants become literals   2 MUFU sin/cos become FMULs   4 no operand-tile stores   8 no TMEM loads
//   16 idle reducers              32 no bulk stores to HBM
#ifndef NVP_ABL
#define NVP_ABL 0
#endif

#if NVP_ABL & 4
#define F_STORE(...) do {} while (0)
#else
#define F_STORE(...) do { __VA_ARGS__; } while (0)
#endif
#if NVP_ABL & 8
#define F_TLOAD(...) do {} while (0)
#else
#define F_TLOAD(...) do { __VA_ARGS__; } while (0)
#endif
#if NVP_ABL & 32
#define F_BULK(...) do {} while (0)
#else
#define F_BULK(...) do { __VA_ARGS__; } while (0)
#endif
#if NVP_ABL & 1
#define KC(x) 0.37f
#else
#define KC(x) (x)
#endif

struct FusedSmem { uint32_t bufs, ring, consts, rowdata, xch, bars, tmem, total; };
__host__ __device__ constexpr FusedSmem fused_smem_layout() {
  FusedSmem s{};
  uint32_t o = 0;
  s.bufs = o; o += 5u * kFBuf;
  s.ring = o; o += kFStages * kPanelBytes;
  s.consts = o; o += (sizeof(FusedConsts) + 15) / 16 * 16;
  s.rowdata = o; o += kTile * 16;        // per row: drgb0..2 (times gs) and tau, each as an fp16 pair
  s.xch = o; o += kFSub * kTile * 16;    // partial rgb of the column slices
  s.bars = o; o += 64 * 8;
  s.tmem = o; o += 16;
  s.total = o;
  return s;
}

#if NVP_ABL & 2
__device__ __forceinline__ float f_sin(float x) { return x * 0.5f; }
__device__ __forceinline__ void f_sincos(float x, float& s, float& c) { s = x * 0.5f; c = x * 0.25f; }
#elif defined(NVP_FUSED_REDUCED_SIN)
__device__ __forceinline__ float f_sin(float x) { return fast_sin<true>(x); }
__device__ __forceinline__ void f_sincos(float x, float& s, float& c) { fast_sincos(x, s, c); }
#else
// sin.approx / cos.approx: FMUL.RZ by 1/2pi + MUFU, which takes its argument in revolutions and drops the integer part
// itself; the only loss against a Cody-Waite reduction is the rounding of that product: measured max error 1.4e-7 |x|
// (8.6e-6 at the |30 (w t + b)| <= 60 of SIREN layer 0; scripts/sin_probe.cu, profiles/r02_sin_probe.txt).
__device__ __forceinline__ float f_sin(float x) { return __sinf(x); }
__device__ __forceinline__ void f_sincos(float x, float& s, float& c) { __sincosf(x, &s, &c); }
#endif

template <int N>
__device__ __forceinline__ void load_rown(const uint8_t* panel, int r, int c0, float (&v)[N]) {
#pragma unroll
  for (int j = 0; j < N / 8; ++j) {
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
__device__ __forceinline__ float2 unpack_half2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
template <int N>
__device__ __forceinline__ void store_packed(uint8_t* panel, int r, int c0, const uint32_t (&v)[N / 2]) {
#pragma unroll
  for (int j = 0; j < N / 8; ++j)
    *reinterpret_cast<uint4*>(panel + panel_chunk_offset(r, (c0 >> 3) + j)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(kFThreads, 1) mlp_fused_kernel(const FusedArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr FusedSmem L = fused_smem_layout();
  constexpr int C = kFCols;
  uint8_t* bufs = smem + L.bufs;
  uint8_t* ring = smem + L.ring;
#if NVP_FCONST == 1
  const FusedConsts& K = g_fused_consts[a.cslot];
#else
  const FusedConsts& K = *reinterpret_cast<const FusedConsts*>(smem + L.consts);
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(FusedConsts) / 4); i += kFThreads)
    reinterpret_cast<float*>(smem + L.consts)[i] = __ldg(a.consts + i);
#endif
  uint4* s_row = reinterpret_cast<uint4*>(smem + L.rowdata);   // per row: (d0,d0) (d1,d1) (d2,d2) (tau,tau) as fp16 pairs
  float4* s_xch = reinterpret_cast<float4*>(smem + L.xch);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L.tmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < FB_COUNT; ++i) {
      uint32_t cnt = 1;
      if (i == FB_TF_P0 || i == FB_TF_P4 || (i >= FB_PD_P0 && i <= FB_PD_P6)) cnt = kFEpiWarps;
      if (i == FB_A2_READY) cnt = kTile;
      if (i >= FB_RD_A2 && i <= FB_RD_DSP0) cnt = 2;
      mbar_init(&bars[i], cnt);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(s_tmem, 512); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t S0 = tmem, S1 = tmem + 128, S2 = tmem + 256, S3 = tmem + 384;

  const int my_tiles = (a.n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  // Shared-memory tiles.  Z and A1 swap their physical buffers from one tile to the next: the next tile's latent is
  // loaded into this tile's A1 buffer (free once G2 has read dm2), so its load overlaps this tile's backward.
  uint8_t* const bH0 = bufs + 1u * kFBuf;
  uint8_t* const bA0 = bufs + 2u * kFBuf;
  uint8_t* const bH1 = bufs + 3u * kFBuf;
  auto buf_z = [&](int it) { return bufs + ((it & 1) ? 4u : 0u) * kFBuf; };
  auto buf_a1 = [&](int it) { return bufs + ((it & 1) ? 0u : 4u) * kFBuf; };
  auto tile_of = [&](int it) { return static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x); };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t g = 0;
      auto issue_w = [&](int idx) {
        const uint32_t st = g % kFStages, ph = (g / kFStages) & 1;
        mbar_wait(&bars[FB_WEMPTY + st], ph ^ 1);
        mbar_arrive_expect_tx(&bars[FB_WFULL + st], kPanelBytes);
        bulk_g2s(ring + st * kPanelBytes, a.wpk + static_cast<size_t>(idx) * kPanelBytes, kPanelBytes, &bars[FB_WFULL + st]);
        ++g;
      };
      mbar_arrive_expect_tx(&bars[FB_ZFULL], kFBuf);
      bulk_g2s(buf_z(0), a.z16t + static_cast<size_t>(tile_of(0)) * kFBuf, kFBuf, &bars[FB_ZFULL]);
      issue_w(25); issue_w(26);   // W0z of the first tile's F0
      for (int it = 0; it < my_tiles; ++it) {
        const bool has_next = it + 1 < my_tiles;
        for (int i = 0; i < kFPanelsPerTile; ++i) {
          if ((i == 25 || i == 26) && !has_next) continue;
          issue_w(i);
          if (i == 21 && has_next) {
            // the A1 buffer (dm2) has been read by G2 and by its bulk store: load the next tile's latent into it
            mbar_wait(&bars[FB_G2_DONE], it & 1);
            mbar_wait(&bars[FB_SR_P3], it & 1);
            mbar_arrive_expect_tx(&bars[FB_ZFULL], kFBuf);
            bulk_g2s(buf_a1(it), a.z16t + static_cast<size_t>(tile_of(it + 1)) * kFBuf, kFBuf, &bars[FB_ZFULL]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(kTile, H, false, false);
      uint32_t g = 0;
#ifdef NVP_TIMELINE
      long long ring_wait = 0;
#endif
      // one 64-wide K panel: A = panel `ap` of a shared-memory tile, B = next ring stage, D = TMEM slot `acc`
      auto gemm = [&](const uint8_t* ap, uint32_t acc, bool accumulate) {
        const uint32_t st = g % kFStages, ph = (g / kFStages) & 1;
#ifdef NVP_TIMELINE
        const long long t0 = clock64();
#endif
        mbar_wait(&bars[FB_WFULL + st], ph);
#ifdef NVP_TIMELINE
        ring_wait += clock64() - t0;
#endif
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(ap), b_addr = smem_u32(ring + st * kPanelBytes);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          umma_f16_ss(acc, umma_desc_kmajor(a_addr, kk), umma_desc_kmajor(b_addr, kk), idesc, accumulate ? 1u : 0u);
          accumulate = true;
        }
        umma_commit(&bars[FB_WEMPTY + st]);
        ++g;
      };
      auto wait_pd = [&](int bar, uint32_t par) { mbar_wait(&bars[bar], par); tcgen05_fence_after(); };
      // This thread also hands finished operand tiles to the TMA for the weight-gradient kernel: it is the one that
      // observes every panel-done barrier anyway, and issuing a cp.async.bulk takes ~150 cycles that the epilogue warps
      // would otherwise spend on their critical path.  Stores are issued after the MMAs of the same panel.
      auto store2 = [&](uint8_t* dst0, const uint8_t* src0, uint8_t* dst1, const uint8_t* src1, uint32_t bytes) {
        F_BULK(bulk_s2g(dst0, src0, bytes));
        F_BULK(if (dst1) bulk_s2g(dst1, src1, bytes));
        bulk_commit();
      };
      // all bulk stores issued so far have left shared memory: the tiles of phase `sr_bar` may be overwritten
      auto stores_read = [&](int sr_bar) { bulk_wait_read0(); mbar_arrive(&bars[sr_bar]); };
      // F0 of the first tile
      wait_pd(FB_ZFULL, 0);
      gemm(buf_z(0), S1, false); gemm(buf_z(0) + kPanelBytes, S1, true);
      umma_commit(&bars[FB_AF_P0]);
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t par = it & 1;
        const bool has_next = it + 1 < my_tiles;
        uint8_t* const bZ = buf_z(it);
        uint8_t* const bA1 = buf_a1(it);
        const uint32_t P = kPanelBytes;
        const int tile = tile_of(it);
        uint8_t* const st_base = a.stash + static_cast<size_t>(tile) * FS_COUNT * kFBuf;
        uint8_t* const dp_base = a.dpre + static_cast<size_t>(tile) * DP_COUNT * kFBuf;
        auto st_dst = [&](int slot, int p) { return st_base + static_cast<size_t>(slot) * kFBuf + static_cast<size_t>(p) * kPanelBytes; };
        auto dp_dst = [&](int slot, int p) { return dp_base + static_cast<size_t>(slot) * kFBuf + static_cast<size_t>(p) * kPanelBytes; };
#ifdef NVP_TIMELINE
        if (it == 3) ring_wait = 0;
#endif
        // ---- F1: m1 -> S0, sp1 -> S1 ----
        NVP_TL(it == 3, 64);
        gemm(bZ, S0, false); gemm(bZ + P, S0, true);
        wait_pd(FB_PD_P0 + 0, par); wait_pd(FB_TF_P0, par);
        gemm(bH0, S0, true); gemm(bA0, S1, false);
        wait_pd(FB_PD_P0 + 1, par);
        NVP_TL(it == 3, 65);
        gemm(bH0 + P, S0, true); gemm(bA0 + P, S1, true);
        umma_commit(&bars[FB_AF_P1]);
        // (stores only after the phase's critical MMAs: a bulk store issued earlier would sit in the SM's TMA queue ahead
        // of the weight panels those MMAs wait for)
        store2(st_dst(FS_H0, 0), bH0, st_dst(FS_A0, 0), bA0, kFBuf);
        NVP_TL(it == 3, 80);
        // ---- F2: m2 -> S2, sp2 -> S3 ----
        gemm(bZ, S2, false); gemm(bZ + P, S2, true);
        stores_read(FB_SR_P0);
        wait_pd(FB_PD_P1 + 0, par);
        gemm(bH1, S2, true); gemm(bA1, S3, false);
        wait_pd(FB_PD_P1 + 1, par);
        NVP_TL(it == 3, 66);
        gemm(bH1 + P, S2, true); gemm(bA1 + P, S3, true);
        umma_commit(&bars[FB_AF_P2]);
        store2(st_dst(FS_H1, 0), bH1, st_dst(FS_A1, 0), bA1, kFBuf);
        stores_read(FB_SR_P1);   // a1 / h1 are on their way: P3 may write dsp2 / dm2 over a0 / a1 (idle time: G2 waits for P2 + P3)
        NVP_TL(it == 3, 81);
        // ---- G2: da1 = dsp2 Ws2 -> S0, dh1 = dm2 W2h -> S2 (dsp2 in A0, dm2 in A1); then dz = dm2 W2z -> S3 ----
        wait_pd(FB_PD_P3 + 0, par);
        NVP_TL(it == 3, 67);
        gemm(bA0, S0, false); gemm(bA1, S2, false);
        wait_pd(FB_PD_P3 + 1, par);
        NVP_TL(it == 3, 68);
        gemm(bA0 + P, S0, true); gemm(bA1 + P, S2, true);
        umma_commit(&bars[FB_AF_P4]);
        NVP_TL(it == 3, 82);
        gemm(bA1, S3, false); gemm(bA1 + P, S3, true);
        umma_commit(&bars[FB_G2_DONE]);
        store2(dp_dst(DP_S2, 0), bA0, dp_dst(DP_M2, 0), bA1, kFBuf);
        stores_read(FB_SR_P3);   // the producer may load the next latent tile over dm2; P5 may write dsp0 over dsp2
        // ---- G1: da0 = dsp1 Ws1 -> S0, dh0 = dm1 W1h -> S2 (dsp1 in Z, dm1 in H1); then dz += dm1 W1z ----
        // S0 / S2 are still read by P4 until it has loaded its last panel from TMEM (FB_TF_P4)
        wait_pd(FB_PD_P4 + 0, par); wait_pd(FB_TF_P4, par);
        NVP_TL(it == 3, 69);
        gemm(bZ, S0, false); gemm(bH1, S2, false);
        wait_pd(FB_PD_P4 + 1, par);
        NVP_TL(it == 3, 70);
        gemm(bZ + P, S0, true); gemm(bH1 + P, S2, true);
        umma_commit(&bars[FB_AF_P5]);
        NVP_TL(it == 3, 83);
        gemm(bH1, S3, true); gemm(bH1 + P, S3, true);
        store2(dp_dst(DP_S1, 0), bZ, dp_dst(DP_M1, 0), bH1, kFBuf);
        stores_read(FB_SR_P4);   // P6 may stage dz over dsp1
        // ---- G0 (dz += dm0 W0z, dm0 in H0) around F0 of the next tile (m0 -> S1, free since FB_TF_P4) ----
        wait_pd(FB_PD_P5 + 0, par);
        NVP_TL(it == 3, 71);
        gemm(bH0, S3, true);
        if (has_next) {
          wait_pd(FB_ZFULL, par ^ 1);
          gemm(bA1, S1, false); gemm(bA1 + P, S1, true);
          umma_commit(&bars[FB_AF_P0]);
        }
        wait_pd(FB_PD_P5 + 1, par);
        NVP_TL(it == 3, 72);
        gemm(bH0 + P, S3, true);
        umma_commit(&bars[FB_AF_P6]);
        store2(dp_dst(DP_M0, 0), bH0, nullptr, nullptr, kFBuf);
        stores_read(FB_SR_P5);   // the next tile's P0 may write h0 over dm0
        // dz, staged in Z by P6
        mbar_wait(&bars[FB_PD_P6], par);
        store2(a.dz16t + static_cast<size_t>(tile) * kFBuf, bZ, nullptr, nullptr, kFBuf);
        stores_read(FB_SR_P6);   // the next tile's P1 may write a1 there
#ifdef NVP_TIMELINE
        if (it == 3 && blockIdx.x == 0) g_timeline[90] = static_cast<unsigned long long>(ring_wait);
#endif
      }
      bulk_wait_all0();
    }
  } else if (warp < 4) {
    // ================= reducers: column sums over the samples of a tile =================
    // Warp w owns the 64 columns of panel w of every operand tile; lane l owns columns 2l, 2l+1 of that panel.  Groups of
    // 8 rows are summed as packed fp16 pairs (the operands are fp16 already), the group sums are accumulated in fp32.
    const int w = warp - 2;
    const uint32_t poff = static_cast<uint32_t>(w) * kPanelBytes;
    const uint32_t lane_off = static_cast<uint32_t>(lane & 3) * 4u;
    const uint32_t lane_chunk = static_cast<uint32_t>(lane >> 2);
    float wl0a = 0.f, wl0b = 0.f, wl1a = 0.f, wl1b = 0.f, wl2a = 0.f, wl2b = 0.f;
    float b2a = 0.f, b2b = 0.f, b1a = 0.f, b1b = 0.f, b0a = 0.f, b0b = 0.f, w0a = 0.f, w0b = 0.f;
    auto ldx = [&](const uint8_t* tile, int r) {
      return *reinterpret_cast<const __half2*>(tile + poff + static_cast<uint32_t>(r) * 128u +
                                               ((lane_chunk ^ (static_cast<uint32_t>(r) & 7u)) << 4) + lane_off);
    };
    auto arrive = [&](int bar) { __syncwarp(); if (lane == 0) mbar_arrive(&bars[bar]); };
    auto colsum = [&](const uint8_t* tile, float& sa, float& sb) {
#if !(NVP_ABL & 16)
#pragma unroll 2
      for (int r0 = 0; r0 < kTile; r0 += 8) {
        __half2 acc = ldx(tile, r0);
#pragma unroll
        for (int j = 1; j < 8; ++j) acc = __hadd2(acc, ldx(tile, r0 + j));
        const float2 f = __half22float2(acc);
        sa += f.x; sb += f.y;
      }
#endif
    };
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t par = it & 1;
      const uint8_t* const bZ = buf_z(it);
      mbar_wait(&bars[FB_A2_READY], par);   // a2 (both panels) and the per-row drgb / tau are in place
#if !(NVP_ABL & 16)
#pragma unroll 2
      for (int r0 = 0; r0 < kTile; r0 += 8) {
        __half2 a0 = __float2half2_rn(0.f), a1 = a0, a2 = a0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __half2 x = ldx(bZ, r0 + j);
          const uint4 d = s_row[r0 + j];   // (drgb0, drgb0) (drgb1, drgb1) (drgb2, drgb2) (tau, tau) as fp16 pairs
          a0 = __hfma2(x, *reinterpret_cast<const __half2*>(&d.x), a0);
          a1 = __hfma2(x, *reinterpret_cast<const __half2*>(&d.y), a1);
          a2 = __hfma2(x, *reinterpret_cast<const __half2*>(&d.z), a2);
        }
        const float2 f0 = __half22float2(a0), f1 = __half22float2(a1), f2 = __half22float2(a2);
        wl0a += f0.x; wl0b += f0.y; wl1a += f1.x; wl1b += f1.y; wl2a += f2.x; wl2b += f2.y;
      }
#endif
      arrive(FB_RD_A2);
      NVP_TL(w == 0 && lane == 0 && it == 3, 48);
      mbar_wait(&bars[FB_PD_P3 + w], par);
      colsum(bA0, b2a, b2b);
      arrive(FB_RD_DSP2);
      NVP_TL(w == 0 && lane == 0 && it == 3, 49);
      mbar_wait(&bars[FB_PD_P4 + w], par);
      colsum(bZ, b1a, b1b);
      arrive(FB_RD_DSP1);
      NVP_TL(w == 0 && lane == 0 && it == 3, 50);
      mbar_wait(&bars[FB_PD_P5 + w], par);
#if !(NVP_ABL & 16)
#pragma unroll 2
      for (int r0 = 0; r0 < kTile; r0 += 8) {
        __half2 a0 = __float2half2_rn(0.f), a1 = a0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __half2 x = ldx(bA0, r0 + j);
          const uint32_t t = s_row[r0 + j].w;
          a0 = __hadd2(a0, x);
          a1 = __hfma2(x, *reinterpret_cast<const __half2*>(&t), a1);
        }
        const float2 f0 = __half22float2(a0), f1 = __half22float2(a1);
        b0a += f0.x; b0b += f0.y; w0a += f1.x; w0b += f1.y;
      }
#endif
      arrive(FB_RD_DSP0);
      NVP_TL(w == 0 && lane == 0 && it == 3, 51);
    }
    if (my_tiles > 0) {
      const float inv_gs = __ldg(a.gscale + 1);
      const int c = w * 64 + 2 * lane;
      if (a.g_last_w) {
        atomicAdd(a.g_last_w + c, wl0a * inv_gs); atomicAdd(a.g_last_w + c + 1, wl0b * inv_gs);
        atomicAdd(a.g_last_w + H + c, wl1a * inv_gs); atomicAdd(a.g_last_w + H + c + 1, wl1b * inv_gs);
        atomicAdd(a.g_last_w + 2 * H + c, wl2a * inv_gs); atomicAdd(a.g_last_w + 2 * H + c + 1, wl2b * inv_gs);
      }
      if (a.g_siren_b2) { atomicAdd(a.g_siren_b2 + c, b2a * inv_gs); atomicAdd(a.g_siren_b2 + c + 1, b2b * inv_gs); }
      if (a.g_siren_b1) { atomicAdd(a.g_siren_b1 + c, b1a * inv_gs); atomicAdd(a.g_siren_b1 + c + 1, b1b * inv_gs); }
      // layer 0's pre-activation is w0 (w t + b): the stored dsp0 omits that factor
      const float s0 = inv_gs * a.w0;
      if (a.g_siren_b0) { atomicAdd(a.g_siren_b0 + c, b0a * s0); atomicAdd(a.g_siren_b0 + c + 1, b0b * s0); }
      if (a.g_siren_w0) { atomicAdd(a.g_siren_w0 + c, w0a * s0); atomicAdd(a.g_siren_w0 + c + 1, w0b * s0); }
    }
  } else {
    // ================= epilogue warps =================
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    const int sub = (warp - 4) >> 2;            // which C-column slice of the current 64-column panel
    const int r = quarter * 32 + lane;          // row inside the tile
    const int pc = sub * C;                     // first column inside the panel
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const bool elected = (tid == 128);   // stamps the timeline (NVP_TIMELINE builds)
    (void)elected;
    const float gs = __ldg(a.gscale), inv_gs = __ldg(a.gscale + 1), loss_mult = __ldg(a.gscale + 2);
    float loss_acc = 0.f, gb0 = 0.f, gb1 = 0.f, gb2 = 0.f;

    // rows are shared by the kFSub warps of a TMEM lane quarter: their barrier (rgb exchange)
    auto quarter_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(2 + quarter), "n"(kFSub * 32) : "memory"); };
    // end of a panel: my operand chunks are visible to the async proxy (UMMA reads, bulk stores); one arrival per warp.
    // There is no barrier among the epilogue warps: each runs on into its next panel.
    auto panel_end = [&](int pd_bar) {
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[pd_bar]);
    };
    auto wait_bar = [&](int bar, uint32_t par) { mbar_wait(&bars[bar], par); };
    auto tmem_release = [&](int bar) { tcgen05_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&bars[bar]); };
    auto wait_acc = [&](int af_bar, uint32_t par) { mbar_wait(&bars[af_bar], par); tcgen05_fence_after(); };

    float tau = 0.f;
    {
      const int64_t s0 = static_cast<int64_t>(tile_of(0)) * kTile + r;
      if (my_tiles > 0 && s0 < a.n) tau = __ldg(a.tau + s0);
    }
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t par = it & 1;
      const int tile = tile_of(it);
      const int64_t s = static_cast<int64_t>(tile) * kTile + r;
      const bool valid = s < a.n;
      uint8_t* const bZ = buf_z(it);
      uint8_t* const bA1 = buf_a1(it);
      // ground truth of my row, fetched a few phases before it is needed
      // (raw bytes: nothing depends on them until P2, so the loads stay in flight behind the phases before it)
      uint32_t g0 = 0, g1 = 0, g2 = 0;
      if (valid && a.dout == nullptr) { g0 = __ldg(a.gt + s * 3); g1 = __ldg(a.gt + s * 3 + 1); g2 = __ldg(a.gt + s * 3 + 2); }

      // ---------------- P0: h0 = lrelu(m0 + b), a0 = sin(w0 (w tau + b)) h0 ----------------
      // writes H0 (dm0 of the previous tile: read by G0, done) and A0 (dsp0 of the previous tile: reducers)
      wait_acc(FB_AF_P0, par);
      NVP_TL(elected && it == 3, 0);
      NVP_TL(elected && it == 4, 14);
      if (it > 0) { wait_bar(FB_SR_P5, par ^ 1); wait_bar(FB_RD_DSP0, par ^ 1); }   // dm0 stored; reducers done with dsp0
      NVP_TL(elected && it == 3, 16);
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        const int col = p * 64 + pc;
        uint32_t vm[C] = {};
        F_TLOAD(tmem_ldn<C>(S1 + lane_base + col, vm));
        tmem_ld_wait();
        if (p == 1) tmem_release(FB_TF_P0);   // S1 may now receive sp1
        float hv[C], av[C];
#pragma unroll
        for (int i = 0; i < C; ++i) {
          hv[i] = lrelu(__uint_as_float(vm[i]));   // the bias came with the GEMM
          av[i] = f_sin(fmaf(tau, KC(K.ws0[col + i]), KC(K.bs[0][col + i]))) * hv[i];
        }
        F_STORE(store_rown<C>(bH0 + p * kPanelBytes, r, pc, hv));
        F_STORE(store_rown<C>(bA0 + p * kPanelBytes, r, pc, av));
        panel_end(FB_PD_P0 + p);
      }
      NVP_TL(elected && it == 3, 1);

      // ---------------- P1: h1 = lrelu(m1 + b), a1 = sin(sp1 + b) h1 ----------------
      // writes H1 (dm1 of the previous tile) and A1 (= the previous tile's Z buffer: its dz store must have left)
      wait_acc(FB_AF_P1, par);
      NVP_TL(elected && it == 3, 2);
      if (it > 0) { wait_bar(FB_SR_P4, par ^ 1); wait_bar(FB_SR_P6, par ^ 1); wait_bar(FB_RD_DSP1, par ^ 1); }   // dm1, dz stored; dsp1 reduced
      NVP_TL(elected && it == 3, 17);
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        const int col = p * 64 + pc;
        uint32_t vm[C] = {}, vs[C] = {};
        F_TLOAD(tmem_ldn<C>(S0 + lane_base + col, vm));
        F_TLOAD(tmem_ldn<C>(S1 + lane_base + col, vs));
        tmem_ld_wait();
        float hv[C], av[C];
#pragma unroll
        for (int i = 0; i < C; ++i) {
          hv[i] = lrelu(__uint_as_float(vm[i]));
          av[i] = f_sin(__uint_as_float(vs[i]) + KC(K.bs[1][col + i])) * hv[i];
        }
        F_STORE(store_rown<C>(bH1 + p * kPanelBytes, r, pc, hv));
        F_STORE(store_rown<C>(bA1 + p * kPanelBytes, r, pc, av));
        panel_end(FB_PD_P1 + p);
      }
      NVP_TL(elected && it == 3, 3);

      // ---------------- P2: layer 2 forward.  a2 -> Z (for the reducers); u = h2 cos2 and v = sin2 lrelu'(m2) stay in
      // registers (packed fp16) until drgb is known ----------------
      wait_acc(FB_AF_P2, par);   // all MMAs reading z, a0, a1 have completed
      NVP_TL(elected && it == 3, 4);
      float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
      uint32_t upk[2][C / 2], vpk[2][C / 2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int col = p * 64 + pc;
        uint32_t vm[C] = {}, vs[C] = {};
        F_TLOAD(tmem_ldn<C>(S2 + lane_base + col, vm));
        F_TLOAD(tmem_ldn<C>(S3 + lane_base + col, vs));
        tmem_ld_wait();
        uint32_t apk[C / 2];
#pragma unroll
        for (int i = 0; i < C; i += 2) {
          float a2[2], uu[2], vv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float m = __uint_as_float(vm[i + e]);
            const float hh = lrelu(m);
            float sn, cs;
            f_sincos(__uint_as_float(vs[i + e]) + KC(K.bs[2][col + i + e]), sn, cs);
            a2[e] = sn * hh;
            rgb0 = fmaf(a2[e], KC(K.wl[0][col + i + e]), rgb0);
            rgb1 = fmaf(a2[e], KC(K.wl[1][col + i + e]), rgb1);
            rgb2 = fmaf(a2[e], KC(K.wl[2][col + i + e]), rgb2);
            uu[e] = hh * cs;
            vv[e] = sn * (m > 0.f ? 1.0f : 0.01f);
          }
          apk[i / 2] = pack_half2(a2[0], a2[1]);
          upk[p][i / 2] = pack_half2(uu[0], uu[1]);
          vpk[p][i / 2] = pack_half2(vv[0], vv[1]);
        }
        F_STORE(store_packed<C>(bZ + p * kPanelBytes, r, pc, apk));
      }
      // rgb of my row = the sum over the column slices
      s_xch[sub * kTile + r] = make_float4(rgb0, rgb1, rgb2, 0.f);
      quarter_sync();
#pragma unroll
      for (int q = 1; q < kFSub; ++q) {
        const float4 o = s_xch[((sub + q) % kFSub) * kTile + r];
        rgb0 += o.x; rgb1 += o.y; rgb2 += o.z;
      }
      rgb0 += KC(K.bl[0]); rgb1 += KC(K.bl[1]); rgb2 += KC(K.bl[2]);
      float d0 = 0.f, d1 = 0.f, d2 = 0.f;
      if (valid) {
        if (a.dout != nullptr) {
          d0 = __ldg(a.dout + s * 3) * gs; d1 = __ldg(a.dout + s * 3 + 1) * gs; d2 = __ldg(a.dout + s * 3 + 2) * gs;
        } else {
          // (pins the conversions here: the compiler otherwise hoists them up to the loads and stalls there for the HBM latency)
          asm volatile("" : "+r"(g0), "+r"(g1), "+r"(g2));
          const float e0 = rgb0 - (static_cast<float>(g0) - 127.5f) / 127.5f;
          const float e1 = rgb1 - (static_cast<float>(g1) - 127.5f) / 127.5f;
          const float e2 = rgb2 - (static_cast<float>(g2) - 127.5f) / 127.5f;
          if (sub == 0) loss_acc += e0 * e0 + e1 * e1 + e2 * e2;
          d0 = e0 * loss_mult; d1 = e1 * loss_mult; d2 = e2 * loss_mult;
        }
        if (sub == 0 && a.rgb_out != nullptr) { a.rgb_out[s * 3] = rgb0; a.rgb_out[s * 3 + 1] = rgb1; a.rgb_out[s * 3 + 2] = rgb2; }
      }
      if (sub == 0) {
        gb0 += d0; gb1 += d1; gb2 += d2;
        s_row[r] = make_uint4(pack_half2(d0, d0), pack_half2(d1, d1), pack_half2(d2, d2), pack_half2(tau, tau));
        mbar_arrive(&bars[FB_A2_READY]);    // releases my a2 / row data (and, through the barrier above, everybody's a2)
      }
      NVP_TL(elected && it == 3, 5);

      // ---------------- P3: da2 = drgb Wl ; dsp2 = da2 u -> A0 ; dm2 = da2 v -> A1 ----------------
      wait_bar(FB_SR_P0, par); wait_bar(FB_SR_P1, par);   // the a0 / a1 stores have left A0 / A1
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int col = p * 64 + pc;
        uint32_t o1[C / 2], o2[C / 2];
#pragma unroll
        for (int i = 0; i < C; i += 2) {
          const float da0 = d0 * KC(K.wl[0][col + i]) + d1 * KC(K.wl[1][col + i]) + d2 * KC(K.wl[2][col + i]);
          const float da1 = d0 * KC(K.wl[0][col + i + 1]) + d1 * KC(K.wl[1][col + i + 1]) + d2 * KC(K.wl[2][col + i + 1]);
          const float2 uu = unpack_half2(upk[p][i / 2]), vv = unpack_half2(vpk[p][i / 2]);
          o1[i / 2] = pack_half2(da0 * uu.x, da1 * uu.y);
          o2[i / 2] = pack_half2(da0 * vv.x, da1 * vv.y);
        }
        F_STORE(store_packed<C>(bA0 + p * kPanelBytes, r, pc, o1));
        F_STORE(store_packed<C>(bA1 + p * kPanelBytes, r, pc, o2));
        panel_end(FB_PD_P3 + p);
      }
      NVP_TL(elected && it == 3, 6);

      // ---------------- P4: dsp1 = da1 h1 cos1 (Z, over a2) ; dm1 = (dh1 + da1 sin1) lrelu'(h1) (H1, in place) ----------------
      wait_acc(FB_AF_P4, par);
      NVP_TL(elected && it == 3, 8);
      wait_bar(FB_RD_A2, par);   // the reducers are done with a2 (Z); h1's store left H1 before P3
      NVP_TL(elected && it == 3, 18);
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        const int col = p * 64 + pc;
        uint32_t va[C] = {}, vh[C] = {}, vs[C] = {};
        F_TLOAD(tmem_ldn<C>(S0 + lane_base + col, va));
        F_TLOAD(tmem_ldn<C>(S2 + lane_base + col, vh));
        F_TLOAD(tmem_ldn<C>(S1 + lane_base + col, vs));
        float hv[C];
        load_rown<C>(bH1 + p * kPanelBytes, r, pc, hv);
        tmem_ld_wait();
        if (p == 1) tmem_release(FB_TF_P4);   // S0 / S2 / S1 may be overwritten (G1, next tile's F0)
        uint32_t o1[C / 2], o2[C / 2];
#pragma unroll
        for (int i = 0; i < C; i += 2) {
          float x[2], y[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float sn, cs;
            f_sincos(__uint_as_float(vs[i + e]) + KC(K.bs[1][col + i + e]), sn, cs);
            const float dav = __uint_as_float(va[i + e]);
            x[e] = dav * hv[i + e] * cs;
            y[e] = fmaf(dav, sn, __uint_as_float(vh[i + e])) * (hv[i + e] > 0.f ? 1.0f : 0.01f);
          }
          o1[i / 2] = pack_half2(x[0], x[1]);
          o2[i / 2] = pack_half2(y[0], y[1]);
        }
        F_STORE(store_packed<C>(bZ + p * kPanelBytes, r, pc, o1));
        F_STORE(store_packed<C>(bH1 + p * kPanelBytes, r, pc, o2));
        panel_end(FB_PD_P4 + p);
      }
      NVP_TL(elected && it == 3, 9);

      // ---------------- P5: dsp0 = da0 h0 cos0 (A0, over dsp2) ; dm0 = (dh0 + da0 sin0) lrelu'(h0) (H0, in place) ----------------
      wait_acc(FB_AF_P5, par);
      NVP_TL(elected && it == 3, 10);
      wait_bar(FB_SR_P3, par); wait_bar(FB_RD_DSP2, par);   // dsp2 stored and reduced (A0); h0's store left H0 before P3
      NVP_TL(elected && it == 3, 19);
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        const int col = p * 64 + pc;
        uint32_t va[C] = {}, vh[C] = {};
        F_TLOAD(tmem_ldn<C>(S0 + lane_base + col, va));
        F_TLOAD(tmem_ldn<C>(S2 + lane_base + col, vh));
        float hv[C];
        load_rown<C>(bH0 + p * kPanelBytes, r, pc, hv);
        tmem_ld_wait();
        uint32_t o1[C / 2], o2[C / 2];
#pragma unroll
        for (int i = 0; i < C; i += 2) {
          float x[2], y[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float sn, cs;
            f_sincos(fmaf(tau, KC(K.ws0[col + i + e]), KC(K.bs[0][col + i + e])), sn, cs);
            const float dav = __uint_as_float(va[i + e]);
            x[e] = dav * hv[i + e] * cs;
            y[e] = fmaf(dav, sn, __uint_as_float(vh[i + e])) * (hv[i + e] > 0.f ? 1.0f : 0.01f);
          }
          o1[i / 2] = pack_half2(x[0], x[1]);
          o2[i / 2] = pack_half2(y[0], y[1]);
        }
        F_STORE(store_packed<C>(bA0 + p * kPanelBytes, r, pc, o1));
        F_STORE(store_packed<C>(bH0 + p * kPanelBytes, r, pc, o2));
        panel_end(FB_PD_P5 + p);
      }
      NVP_TL(elected && it == 3, 11);
      // tau of the next tile's row (consumed at its P0, needed by the reducers until this tile's dsp0 sum is done:
      // it only enters s_row at the next tile's P2)
      {
        const int64_t sn = s + static_cast<int64_t>(gridDim.x) * kTile;
        tau = (it + 1 < my_tiles && sn < a.n) ? __ldg(a.tau + sn) : 0.0f;
      }

      // ---------------- P6: dz -> fp16 tile staged in Z (over dsp1) -> HBM ----------------
      wait_acc(FB_AF_P6, par);
      NVP_TL(elected && it == 3, 12);
      wait_bar(FB_SR_P4, par); wait_bar(FB_RD_DSP1, par);   // dsp1 stored and reduced (Z)
      NVP_TL(elected && it == 3, 20);
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        uint32_t v[C] = {};
        float f[C];
        F_TLOAD(tmem_ldn<C>(S3 + lane_base + q * 64 + pc, v));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < C; ++i) f[i] = __uint_as_float(v[i]);
        F_STORE(store_rown<C>(bZ + q * kPanelBytes, r, pc, f));
      }
      panel_end(FB_PD_P6);   // the MMA thread stores the tile
      NVP_TL(elected && it == 3, 13);
    }
    if (my_tiles > 0 && sub == 0) {
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) {
        loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o2);
        gb0 += __shfl_xor_sync(0xffffffffu, gb0, o2);
        gb1 += __shfl_xor_sync(0xffffffffu, gb1, o2);
        gb2 += __shfl_xor_sync(0xffffffffu, gb2, o2);
      }
      if (lane == 0) {
        if (a.loss_sum && a.dout == nullptr) atomicAdd(a.loss_sum, loss_acc);
        if (a.g_last_b) {
          atomicAdd(a.g_last_b, gb0 * inv_gs); atomicAdd(a.g_last_b + 1, gb1 * inv_gs); atomicAdd(a.g_last_b + 2, gb2 * inv_gs);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// Weight stream of the fused kernel, one 16 KiB panel per ring stage, in the MMA warp's consumption order:
//   F1  [W1z 0][W1z 1][W1h 0][Ws1 0][W1h 1][Ws1 1]        F2  [W2z 0][W2z 1][W2h 0][Ws2 0][W2h 1][Ws2 1]
//   G2  [Ws2^T 0][W2h^T 0][Ws2^T 1][W2h^T 1] (da1, dh1: the next phase waits for them) [W2z^T 0][W2z^T 1] (dz)
//   G1  the same for layer 1
//   G0 / next F0  [W0z^T 0][W0z 0][W0z 1][W0z^T 1]
// Forward panels: 64 K columns of a [128 out x K] matrix; backward panels: 64 output features (K) of the transposed
// [N = input features, 128] matrix.
int pack_fused_weights(const nvp_desc* d, const nvp_params* p, uint8_t* dst, cudaStream_t st) {
  const Dims m = make_dims(d);
  PackArgs a{};
  a.dst = dst;
  uint32_t off = 0;
  // forward latent panels carry the modulator bias in the latent's two constant-1 columns Z, Z+1 (hi + lo fp16 parts)
  auto fwd_z = [&](int i, int q) {
    if (i == 0) add_panel(a, p->mod_w[0], m.Z, 0, 0, 64 * q, H, m.Z - 64 * q, H, off);
    else add_panel(a, p->mod_w[i], H + m.Z, 0, 0, H + 64 * q, H, m.Z - 64 * q, H, off);
    PackPanel& pp = a.p[a.n - 1];
    pp.bias = p->mod_b[i];
    const int hi = m.Z - 64 * q, lo = m.Z + 1 - 64 * q;
    pp.bias_hi_c = (hi >= 0 && hi < 64) ? hi : -1;
    pp.bias_lo_c = (lo >= 0 && lo < 64) ? lo : -1;
  };
  auto fwd_h = [&](int i, int q) { add_panel(a, p->mod_w[i], H + m.Z, 0, 0, 64 * q, H, 64, H, off); };
  auto fwd_s = [&](int i, int q) { add_panel(a, p->siren_w[i], H, 0, 0, 64 * q, H, 64, H, off); };
  auto bwd_z = [&](int i, int q) {
    if (i == 0) add_panel(a, p->mod_w[0], m.Z, 1, 0, 64 * q, m.Z, 64, H, off);
    else add_panel(a, p->mod_w[i], H + m.Z, 1, H, 64 * q, m.Z, 64, H, off);
  };
  auto bwd_h = [&](int i, int q) { add_panel(a, p->mod_w[i], H + m.Z, 1, 0, 64 * q, H, 64, H, off); };
  auto bwd_s = [&](int i, int q) { add_panel(a, p->siren_w[i], H, 1, 0, 64 * q, H, 64, H, off); };
  for (int i = 1; i <= 2; ++i) {
    fwd_z(i, 0); fwd_z(i, 1);
    for (int q = 0; q < 2; ++q) { fwd_h(i, q); fwd_s(i, q); }
  }
  for (int i = 2; i >= 1; --i) {
    for (int q = 0; q < 2; ++q) { bwd_s(i, q); bwd_h(i, q); }
    bwd_z(i, 0); bwd_z(i, 1);
  }
  bwd_z(0, 0); fwd_z(0, 0); fwd_z(0, 1); bwd_z(0, 1);
  ScopedKernelTimer timer(K_PACK, st);
  pack_weights_kernel<<<dim3(a.n, 4), 256, 0, st>>>(a);
  NVP_LAUNCH_CHECK();
  return 0;
}
