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
// One persistent CTA per SM, one tile in flight, 12 warps:
//   warp 0       TMA producer: weight panel ring (28 panels per tile, in MMA consumption order) + latent tile
//   warp 1       MMA issuer (one thread): tcgen05.mma M=128, N=128, K=16, accumulators in four 128-column TMEM slots
//   warps 2-3    reducers: column sums over the tile's samples (dW_last, SIREN bias gradients, layer-0 w/b gradients)
//                read from the fp16 operand tiles in shared memory, accumulated in fp32 registers across tiles
//   warps 4-11   epilogue: tcgen05.ld -> fp32 math -> fp16 operand tiles written in the UMMA layout
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

constexpr int kFThreads = 384;
constexpr int kFEpiThreads = 256;
constexpr int kFStages = 3;
constexpr int kFPanelsPerTile = 28;
constexpr uint32_t kFBuf = 2u * kPanelBytes;   // one [128 x 128] fp16 operand tile
enum { FS_H0 = 0, FS_A0, FS_H1, FS_A1, FS_COUNT };   // stash slots the weight-gradient kernel reads

// barrier slots
enum {
  FB_WFULL = 0, FB_WEMPTY = 3, FB_ZFULL = 6,
  FB_AF_P0 = 7, FB_AF_P1, FB_AF_P2, FB_AF_P4, FB_AF_P5, FB_AF_P6,
  FB_PD_P0 = 13, FB_PD_P1 = 15, FB_PD_P3 = 17, FB_PD_P4 = 19, FB_PD_P5 = 21,
  FB_TF_P0 = 23, FB_TF_P4, FB_G2_DONE, FB_DM2_STORED, FB_RD_A2, FB_RD_DSP2, FB_RD_DSP1, FB_RD_DSP0, FB_COUNT
};

struct FusedArgs {
  const uint8_t* wpk;      // 28 weight panels in stream order (see pack_fused_weights)
  const uint8_t* z16t;     // latent tiles [tile][2 panels]
  const float* tau;
  const float* mod_b[3];
  const float* siren_b[3];
  const float* siren_w0;
  const float* last_w;
  const float* last_b;
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

struct FusedSmem { uint32_t bufs, ring, consts, rowdata, xch, bars, tmem, total; };
__host__ __device__ constexpr FusedSmem fused_smem_layout() {
  FusedSmem s{};
  uint32_t o = 0;
  s.bufs = o; o += 5u * kFBuf;
  s.ring = o; o += kFStages * kPanelBytes;
  s.consts = o; o += (10 * H + 4) * 4;   // bm[3][H] bs[3][H] (layer 0 times w0) ws0[H] (times w0) wl[3][H] bl[3]
  s.rowdata = o; o += kTile * 16;        // per row: drgb0..2 (times gs), tau
  s.xch = o; o += 2 * kTile * 16;        // partial rgb of the two column halves
  s.bars = o; o += 64 * 8;
  s.tmem = o; o += 16;
  s.total = o;
  return s;
}

#ifdef NVP_FUSED_REDUCED_SIN
__device__ __forceinline__ float f_sin(float x) { return fast_sin<true>(x); }
__device__ __forceinline__ void f_sincos(float x, float& s, float& c) { fast_sincos(x, s, c); }
#else
// sin.approx / cos.approx: FMUL.RZ by 1/2pi + MUFU, which takes its argument in revolutions and drops the integer part
// itself; the only loss against a Cody-Waite reduction is the rounding of that product (|x| * 2^-24, i.e. 4e-6 at the
// |30 (w t + b)| <= 60 of SIREN layer 0; scripts/sin_probe.cu measures it).
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

__global__ void __launch_bounds__(kFThreads, 1) mlp_fused_kernel(const FusedArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr FusedSmem L = fused_smem_layout();
  uint8_t* bufs = smem + L.bufs;
  uint8_t* ring = smem + L.ring;
  float* s_bm = reinterpret_cast<float*>(smem + L.consts);   // [3][H]
  float* s_bs = s_bm + 3 * H;                                // [3][H], layer 0 pre-multiplied by w0
  float* s_ws0 = s_bm + 6 * H;                               // [H], pre-multiplied by w0
  float* s_wl = s_bm + 7 * H;                                // [3][H]
  float* s_bl = s_bm + 10 * H;                               // [3]
  float4* s_row = reinterpret_cast<float4*>(smem + L.rowdata);
  float4* s_xch = reinterpret_cast<float4*>(smem + L.xch);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L.tmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 3 * H; i += kFThreads) {
    s_bm[i] = __ldg(a.mod_b[i / H] + (i % H));
    s_bs[i] = __ldg(a.siren_b[i / H] + (i % H)) * (i < H ? a.w0 : 1.0f);
    s_wl[i] = __ldg(a.last_w + i);
  }
  for (int i = tid; i < H; i += kFThreads) s_ws0[i] = __ldg(a.siren_w0 + i) * a.w0;
  if (tid < 3) s_bl[tid] = __ldg(a.last_b + tid);
  if (tid == 0) {
    for (int i = 0; i < FB_COUNT; ++i) {
      uint32_t cnt = 1;
      if (i == FB_TF_P0 || i == FB_TF_P4) cnt = kFEpiThreads / 32;
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
          if (i == 20 && has_next) {
            // the A1 buffer (dm2) has been read by G2 and by its bulk store: load the next tile's latent into it
            mbar_wait(&bars[FB_G2_DONE], it & 1);
            mbar_wait(&bars[FB_DM2_STORED], it & 1);
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
      // one 64-wide K panel: A = panel `ap` of a shared-memory tile, B = next ring stage, D = TMEM slot `acc`
      auto gemm = [&](const uint8_t* ap, uint32_t acc, bool accumulate) {
        const uint32_t st = g % kFStages, ph = (g / kFStages) & 1;
        mbar_wait(&bars[FB_WFULL + st], ph);
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
        // ---- F1: m1 -> S0, sp1 -> S1 ----
        NVP_TL(it == 3, 64);
        gemm(bZ, S0, false); gemm(bZ + P, S0, true);
        wait_pd(FB_PD_P0 + 0, par); wait_pd(FB_TF_P0, par);
        gemm(bH0, S0, true); gemm(bA0, S1, false);
        wait_pd(FB_PD_P0 + 1, par);
        NVP_TL(it == 3, 65);
        gemm(bH0 + P, S0, true); gemm(bA0 + P, S1, true);
        umma_commit(&bars[FB_AF_P1]);
        // ---- F2: m2 -> S2, sp2 -> S3 ----
        gemm(bZ, S2, false); gemm(bZ + P, S2, true);
        wait_pd(FB_PD_P1 + 0, par);
        gemm(bH1, S2, true); gemm(bA1, S3, false);
        wait_pd(FB_PD_P1 + 1, par);
        NVP_TL(it == 3, 66);
        gemm(bH1 + P, S2, true); gemm(bA1 + P, S3, true);
        umma_commit(&bars[FB_AF_P2]);
        // ---- G2: da1 = dsp2 Ws2 -> S0, dh1 = dm2 W2h -> S2, dz = dm2 W2z -> S3 (dsp2 in A0, dm2 in A1) ----
        wait_pd(FB_PD_P3 + 0, par);
        NVP_TL(it == 3, 67);
        gemm(bA0, S0, false); gemm(bA1, S2, false); gemm(bA1, S3, false);
        wait_pd(FB_PD_P3 + 1, par);
        NVP_TL(it == 3, 68);
        gemm(bA0 + P, S0, true); gemm(bA1 + P, S2, true);
        umma_commit(&bars[FB_AF_P4]);
        gemm(bA1 + P, S3, true);
        umma_commit(&bars[FB_G2_DONE]);
        // ---- G1: da0 = dsp1 Ws1 -> S0, dh0 = dm1 W1h -> S2, dz += dm1 W1z (dsp1 in Z, dm1 in H1) ----
        // S0 / S2 are still read by P4 until it has loaded its last panel from TMEM (FB_TF_P4)
        wait_pd(FB_PD_P4 + 0, par); wait_pd(FB_TF_P4, par);
        NVP_TL(it == 3, 69);
        gemm(bZ, S0, false); gemm(bH1, S2, false); gemm(bH1, S3, true);
        wait_pd(FB_PD_P4 + 1, par);
        NVP_TL(it == 3, 70);
        gemm(bZ + P, S0, true); gemm(bH1 + P, S2, true);
        umma_commit(&bars[FB_AF_P5]);
        gemm(bH1 + P, S3, true);
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
      }
    }
  } else if (warp < 4) {
    // ================= reducers: column sums over the samples of a tile =================
    // Warp w owns the 64 columns of panel w of every operand tile; lane l owns columns 2l, 2l+1 of that panel.
    const int w = warp - 2;
    const uint32_t poff = static_cast<uint32_t>(w) * kPanelBytes;
    const uint32_t lane_off = static_cast<uint32_t>(lane & 3) * 4u;
    const uint32_t lane_chunk = static_cast<uint32_t>(lane >> 2);
    float wl0a = 0.f, wl0b = 0.f, wl1a = 0.f, wl1b = 0.f, wl2a = 0.f, wl2b = 0.f;
    float b2a = 0.f, b2b = 0.f, b1a = 0.f, b1b = 0.f, b0a = 0.f, b0b = 0.f, w0a = 0.f, w0b = 0.f;
    auto ldx = [&](const uint8_t* tile, int r) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(tile + poff + static_cast<uint32_t>(r) * 128u +
                                                            ((lane_chunk ^ (static_cast<uint32_t>(r) & 7u)) << 4) + lane_off);
      return __half22float2(*reinterpret_cast<const __half2*>(&v));
    };
    auto arrive = [&](int bar) { __syncwarp(); if (lane == 0) mbar_arrive(&bars[bar]); };
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t par = it & 1;
      const uint8_t* const bZ = buf_z(it);
      mbar_wait(&bars[FB_PD_P3 + 0], par);   // a2 (both panels), the per-row data and dsp2 panel 0 are in place
#pragma unroll 8
      for (int r = 0; r < kTile; ++r) {
        const float2 x = ldx(bZ, r);
        const float4 d = s_row[r];
        wl0a = fmaf(d.x, x.x, wl0a); wl0b = fmaf(d.x, x.y, wl0b);
        wl1a = fmaf(d.y, x.x, wl1a); wl1b = fmaf(d.y, x.y, wl1b);
        wl2a = fmaf(d.z, x.x, wl2a); wl2b = fmaf(d.z, x.y, wl2b);
      }
      arrive(FB_RD_A2);
      if (w == 1) mbar_wait(&bars[FB_PD_P3 + 1], par);
#pragma unroll 8
      for (int r = 0; r < kTile; ++r) { const float2 x = ldx(bA0, r); b2a += x.x; b2b += x.y; }
      arrive(FB_RD_DSP2);
      mbar_wait(&bars[FB_PD_P4 + w], par);
#pragma unroll 8
      for (int r = 0; r < kTile; ++r) { const float2 x = ldx(bZ, r); b1a += x.x; b1b += x.y; }
      arrive(FB_RD_DSP1);
      mbar_wait(&bars[FB_PD_P5 + w], par);
#pragma unroll 8
      for (int r = 0; r < kTile; ++r) {
        const float2 x = ldx(bA0, r);
        const float t = s_row[r].w;
        b0a += x.x; b0b += x.y;
        w0a = fmaf(t, x.x, w0a); w0b = fmaf(t, x.y, w0b);
      }
      arrive(FB_RD_DSP0);
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
    const int sub = (warp - 4) >> 2;            // which 32-column half of the current 64-column panel
    const int r = quarter * 32 + lane;          // row inside the tile
    const int pc = sub * 32;                    // first column inside the panel
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const bool elected = (tid == 128);
    const float gs = __ldg(a.gscale), inv_gs = __ldg(a.gscale + 1), loss_mult = __ldg(a.gscale + 2);
    float loss_acc = 0.f, gb0 = 0.f, gb1 = 0.f, gb2 = 0.f;

    auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kFEpiThreads) : "memory"); };
    // start of a phase: the accumulators are complete; every earlier bulk store has left its shared-memory source
    // and the reducers are done with the tile this phase overwrites (rd_bar >= 0) before anybody writes
    auto phase_begin = [&](int af_bar, uint32_t par, int rd_bar, uint32_t rd_par, bool rd_wait) {
      mbar_wait(&bars[af_bar], par);
      tcgen05_fence_after();
      if (elected) {
        bulk_wait_read0();
        if (rd_wait) mbar_wait(&bars[rd_bar], rd_par);
      }
      epi_sync();
    };
    // end of a panel: operand tiles visible to the async proxy, then one thread releases the MMA warp and hands the
    // panel(s) to the TMA for the weight-gradient kernel
    auto panel_end = [&](int pd_bar, uint8_t* dst0, const uint8_t* src0, uint8_t* dst1, const uint8_t* src1) {
      fence_proxy_async_smem();
      tcgen05_fence_before();
      epi_sync();
      if (elected) {
        mbar_arrive(&bars[pd_bar]);
        if (dst0) bulk_s2g(dst0, src0, kPanelBytes);
        if (dst1) bulk_s2g(dst1, src1, kPanelBytes);
        bulk_commit();
      }
    };
    auto tmem_release = [&](int bar) { tcgen05_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&bars[bar]); };

    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t par = it & 1;
      const int tile = tile_of(it);
      const int64_t s = static_cast<int64_t>(tile) * kTile + r;
      const bool valid = s < a.n;
      const float tau = valid ? __ldg(a.tau + s) : 0.0f;
      uint8_t* const bZ = buf_z(it);
      uint8_t* const bA1 = buf_a1(it);
      uint8_t* const st_base = a.stash + static_cast<size_t>(tile) * FS_COUNT * kFBuf;
      uint8_t* const dp_base = a.dpre + static_cast<size_t>(tile) * DP_COUNT * kFBuf;
      auto st_dst = [&](int slot, int p) { return st_base + static_cast<size_t>(slot) * kFBuf + static_cast<size_t>(p) * kPanelBytes; };
      auto dp_dst = [&](int slot, int p) { return dp_base + static_cast<size_t>(slot) * kFBuf + static_cast<size_t>(p) * kPanelBytes; };

      // ---------------- P0: h0 = lrelu(m0 + b), a0 = sin(w0 (w tau + b)) h0 ----------------
      // writes H0 (dm0 of the previous tile: read by G0, done) and A0 (dsp0 of the previous tile: reducers)
      phase_begin(FB_AF_P0, par, FB_RD_DSP0, par ^ 1, it > 0);
      NVP_TL(elected && it == 3, 0);
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        const int col = p * 64 + pc;
        uint32_t vm[32];
        tmem_ld32(S1 + lane_base + col, vm);
        tmem_ld_wait();
        if (p == 1) tmem_release(FB_TF_P0);   // S1 may now receive sp1
        float hv[32], av[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          hv[i] = lrelu(__uint_as_float(vm[i]) + s_bm[col + i]);
          av[i] = f_sin(fmaf(tau, s_ws0[col + i], s_bs[col + i])) * hv[i];
        }
        store_row32(bH0 + p * kPanelBytes, r, pc, hv);
        store_row32(bA0 + p * kPanelBytes, r, pc, av);
        panel_end(FB_PD_P0 + p, st_dst(FS_H0, p), bH0 + p * kPanelBytes, st_dst(FS_A0, p), bA0 + p * kPanelBytes);
      }
      NVP_TL(elected && it == 3, 1);

      // ---------------- P1: h1 = lrelu(m1 + b), a1 = sin(sp1 + b) h1 ----------------
      phase_begin(FB_AF_P1, par, 0, 0, false);
      NVP_TL(elected && it == 3, 2);
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        const int col = p * 64 + pc;
        uint32_t vm[32], vs[32];
        tmem_ld32(S0 + lane_base + col, vm);
        tmem_ld32(S1 + lane_base + col, vs);
        tmem_ld_wait();
        float hv[32], av[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          hv[i] = lrelu(__uint_as_float(vm[i]) + s_bm[H + col + i]);
          av[i] = f_sin(__uint_as_float(vs[i]) + s_bs[H + col + i]) * hv[i];
        }
        store_row32(bH1 + p * kPanelBytes, r, pc, hv);
        store_row32(bA1 + p * kPanelBytes, r, pc, av);
        panel_end(FB_PD_P1 + p, st_dst(FS_H1, p), bH1 + p * kPanelBytes, st_dst(FS_A1, p), bA1 + p * kPanelBytes);
      }
      NVP_TL(elected && it == 3, 3);

      // ---------------- P2: layer 2 forward; keeps u = h2 cos2 (A0) and v = sin2 lrelu'(m2) (A1), a2 (Z) ----------------
      phase_begin(FB_AF_P2, par, 0, 0, false);
      NVP_TL(elected && it == 3, 4);
      float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        const int col = p * 64 + pc;
        uint32_t vm[32], vs[32];
        tmem_ld32(S2 + lane_base + col, vm);
        tmem_ld32(S3 + lane_base + col, vs);
        tmem_ld_wait();
        float x0[32], x1[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float m = __uint_as_float(vm[i]) + s_bm[2 * H + col + i];
          const float hh = lrelu(m);
          float sn, cs;
          f_sincos(__uint_as_float(vs[i]) + s_bs[2 * H + col + i], sn, cs);
          const float a2 = sn * hh;
          rgb0 = fmaf(a2, s_wl[col + i], rgb0);
          rgb1 = fmaf(a2, s_wl[H + col + i], rgb1);
          rgb2 = fmaf(a2, s_wl[2 * H + col + i], rgb2);
          x0[i] = a2;
          x1[i] = hh * cs;                                   // u
          vm[i] = __float_as_uint(sn * (m > 0.f ? 1.0f : 0.01f));   // v
        }
        store_row32(bZ + p * kPanelBytes, r, pc, x0);
        store_row32(bA0 + p * kPanelBytes, r, pc, x1);
#pragma unroll
        for (int i = 0; i < 32; ++i) x0[i] = __uint_as_float(vm[i]);
        store_row32(bA1 + p * kPanelBytes, r, pc, x0);
      }
      // rgb of my row = my half of the columns + the other warp's half
      s_xch[sub * kTile + r] = make_float4(rgb0, rgb1, rgb2, 0.f);
      epi_sync();
      {
        const float4 o = s_xch[(sub ^ 1) * kTile + r];
        rgb0 += o.x + s_bl[0]; rgb1 += o.y + s_bl[1]; rgb2 += o.z + s_bl[2];
      }
      float d0 = 0.f, d1 = 0.f, d2 = 0.f;
      if (valid) {
        if (a.dout != nullptr) {
          d0 = __ldg(a.dout + s * 3) * gs; d1 = __ldg(a.dout + s * 3 + 1) * gs; d2 = __ldg(a.dout + s * 3 + 2) * gs;
        } else {
          const float e0 = rgb0 - (static_cast<float>(__ldg(a.gt + s * 3)) - 127.5f) / 127.5f;
          const float e1 = rgb1 - (static_cast<float>(__ldg(a.gt + s * 3 + 1)) - 127.5f) / 127.5f;
          const float e2 = rgb2 - (static_cast<float>(__ldg(a.gt + s * 3 + 2)) - 127.5f) / 127.5f;
          if (sub == 0) loss_acc += e0 * e0 + e1 * e1 + e2 * e2;
          d0 = e0 * loss_mult; d1 = e1 * loss_mult; d2 = e2 * loss_mult;
        }
        if (sub == 0 && a.rgb_out != nullptr) { a.rgb_out[s * 3] = rgb0; a.rgb_out[s * 3 + 1] = rgb1; a.rgb_out[s * 3 + 2] = rgb2; }
      }
      if (sub == 0) {
        gb0 += d0; gb1 += d1; gb2 += d2;
        s_row[r] = make_float4(d0, d1, d2, tau);
      }
      NVP_TL(elected && it == 3, 5);

      // ---------------- P3: da2 = drgb Wl ; dsp2 = da2 u (A0) ; dm2 = da2 v (A1) ----------------
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        const int col = p * 64 + pc;
        float u[32], da[32];
        load_rown<32>(bA0 + p * kPanelBytes, r, pc, u);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          da[i] = d0 * s_wl[col + i] + d1 * s_wl[H + col + i] + d2 * s_wl[2 * H + col + i];
          u[i] *= da[i];
        }
        store_row32(bA0 + p * kPanelBytes, r, pc, u);
        load_rown<32>(bA1 + p * kPanelBytes, r, pc, u);
#pragma unroll
        for (int i = 0; i < 32; ++i) u[i] *= da[i];
        store_row32(bA1 + p * kPanelBytes, r, pc, u);
        panel_end(FB_PD_P3 + p, dp_dst(DP_S2, p), bA0 + p * kPanelBytes, dp_dst(DP_M2, p), bA1 + p * kPanelBytes);
      }
      NVP_TL(elected && it == 3, 6);

      // ---------------- P4: dsp1 = da1 h1 cos1 (Z, over a2) ; dm1 = (dh1 + da1 sin1) lrelu'(h1) (H1, in place) ----------------
      mbar_wait(&bars[FB_AF_P4], par);
      tcgen05_fence_after();
      if (elected) {
        bulk_wait_read0();                        // dsp2 / dm2 have left A0 / A1 ...
        mbar_arrive(&bars[FB_DM2_STORED]);        // ... so the producer may load the next latent tile over dm2
        mbar_wait(&bars[FB_RD_A2], par);          // the reducers are done with a2
      }
      epi_sync();
      NVP_TL(elected && it == 3, 8);
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          const int c16 = pc + ch * 16, col = p * 64 + c16;
          uint32_t va[16], vh[16], vs[16];
          tmem_ld16(S0 + lane_base + col, va);
          tmem_ld16(S2 + lane_base + col, vh);
          tmem_ld16(S1 + lane_base + col, vs);
          float hv[16], o[16];
          load_rown<16>(bH1 + p * kPanelBytes, r, c16, hv);
          tmem_ld_wait();
          if (p == 1 && ch == 1) tmem_release(FB_TF_P4);   // S0 / S2 / S1 may be overwritten (G1, next tile's F0)
          float q[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float sn, cs;
            f_sincos(__uint_as_float(vs[i]) + s_bs[H + col + i], sn, cs);
            const float dav = __uint_as_float(va[i]);
            o[i] = dav * hv[i] * cs;
            q[i] = fmaf(dav, sn, __uint_as_float(vh[i])) * (hv[i] > 0.f ? 1.0f : 0.01f);
          }
          store_rown<16>(bZ + p * kPanelBytes, r, c16, o);
          store_rown<16>(bH1 + p * kPanelBytes, r, c16, q);
        }
        panel_end(FB_PD_P4 + p, dp_dst(DP_S1, p), bZ + p * kPanelBytes, dp_dst(DP_M1, p), bH1 + p * kPanelBytes);
      }
      NVP_TL(elected && it == 3, 9);

      // ---------------- P5: dsp0 = da0 h0 cos0 (A0, over dsp2) ; dm0 = (dh0 + da0 sin0) lrelu'(h0) (H0, in place) ----------------
      phase_begin(FB_AF_P5, par, FB_RD_DSP2, par, true);
      NVP_TL(elected && it == 3, 10);
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          const int c16 = pc + ch * 16, col = p * 64 + c16;
          uint32_t va[16], vh[16];
          tmem_ld16(S0 + lane_base + col, va);
          tmem_ld16(S2 + lane_base + col, vh);
          float hv[16], o[16], q[16];
          load_rown<16>(bH0 + p * kPanelBytes, r, c16, hv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float sn, cs;
            f_sincos(fmaf(tau, s_ws0[col + i], s_bs[col + i]), sn, cs);
            const float dav = __uint_as_float(va[i]);
            o[i] = dav * hv[i] * cs;
            q[i] = fmaf(dav, sn, __uint_as_float(vh[i])) * (hv[i] > 0.f ? 1.0f : 0.01f);
          }
          store_rown<16>(bA0 + p * kPanelBytes, r, c16, o);
          store_rown<16>(bH0 + p * kPanelBytes, r, c16, q);
        }
        panel_end(FB_PD_P5 + p, dp_dst(DP_M0, p), bH0 + p * kPanelBytes, nullptr, nullptr);
      }
      NVP_TL(elected && it == 3, 11);

      // ---------------- P6: dz -> fp16 tile staged in Z (over dsp1) -> HBM ----------------
      phase_begin(FB_AF_P6, par, FB_RD_DSP1, par, true);
      NVP_TL(elected && it == 3, 12);
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        uint32_t v[32];
        float f[32];
        tmem_ld32(S3 + lane_base + q * 64 + pc, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        store_row32(bZ + q * kPanelBytes, r, pc, f);
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      epi_sync();
      if (elected) {
        bulk_s2g(a.dz16t + static_cast<size_t>(tile) * kFBuf, bZ, kFBuf);
        bulk_commit();
      }
      NVP_TL(elected && it == 3, 13);
    }
    if (elected) bulk_wait_all0();
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
//   G2  [Ws2^T 0][W2h^T 0][W2z^T 0][Ws2^T 1][W2h^T 1][W2z^T 1]   G1  the same for layer 1
//   G0 / next F0  [W0z^T 0][W0z 0][W0z 1][W0z^T 1]
// Forward panels: 64 K columns of a [128 out x K] matrix; backward panels: 64 output features (K) of the transposed
// [N = input features, 128] matrix.
int pack_fused_weights(const nvp_desc* d, const nvp_params* p, uint8_t* dst, cudaStream_t st) {
  const Dims m = make_dims(d);
  PackArgs a{};
  a.dst = dst;
  uint32_t off = 0;
  auto fwd_z = [&](int i, int q) {
    if (i == 0) add_panel(a, p->mod_w[0], m.Z, 0, 0, 64 * q, H, m.Z - 64 * q, H, off);
    else add_panel(a, p->mod_w[i], H + m.Z, 0, 0, H + 64 * q, H, m.Z - 64 * q, H, off);
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
  for (int i = 2; i >= 1; --i)
    for (int q = 0; q < 2; ++q) { bwd_s(i, q); bwd_h(i, q); bwd_z(i, q); }
  bwd_z(0, 0); fwd_z(0, 0); fwd_z(0, 1); bwd_z(0, 1);
  ScopedKernelTimer timer(K_PACK, st);
  pack_weights_kernel<<<a.n, 256, 0, st>>>(a);
  NVP_LAUNCH_CHECK();
  return 0;
}
