// Single-tile tcgen05 self-test: one CTA computes a 128x128 fp32 product from fp16 operands staged
// in shared memory in the panel format of tc_common.cuh.  Used by tests/test_gpu_tc_primitives.py to
// check the UMMA descriptors, the manual 128-byte swizzle and the TMEM read-back on real hardware.
//   mode 0 (K-major operands) : D = A[128,K] * B[128,K]^T          K % 64 == 0, K <= 256
//   mode 1 (MN-major operands): D = A[K,128]^T * B[K,128]          K % 16 == 0, K <= 128
//   mode 2+x, x in 0..3 (skinny): D[:, 0:16] = A[K,128]^T * B[K, 16x : 16x+16]   (N = 16, B start offset
//                               inside the 128-byte swizzle atom: the column-sum / small-gradient trick)
#include "common.cuh"
#include "tc_common.cuh"

namespace nvp {
namespace {
using namespace tc;

__global__ void __launch_bounds__(128) umma_selftest_kernel(const __half* __restrict__ A, const __half* __restrict__ B,
                                                            float* __restrict__ D, int K, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int t = threadIdx.x, warp = t >> 5;
  const int npan = mode == 0 ? K / 64 : 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + npan * kPanelBytes;

  if (warp == 0) {
    tmem_alloc(&s_tmem, 128);
    tmem_relinquish();
  }
  if (t == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (mode == 0) {
    for (int p = 0; p < npan; ++p)
      for (int j = 0; j < 8; ++j) {
        const uint4 va = *reinterpret_cast<const uint4*>(A + static_cast<size_t>(t) * K + p * 64 + j * 8);
        const uint4 vb = *reinterpret_cast<const uint4*>(B + static_cast<size_t>(t) * K + p * 64 + j * 8);
        *reinterpret_cast<uint4*>(sA + p * kPanelBytes + panel_chunk_offset(t, j)) = va;
        *reinterpret_cast<uint4*>(sB + p * kPanelBytes + panel_chunk_offset(t, j)) = vb;
      }
  } else {
    for (int q = 0; q < 2; ++q)
      for (int j = 0; j < 8; ++j) {
        uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
        if (t < K) {
          va = *reinterpret_cast<const uint4*>(A + static_cast<size_t>(t) * 128 + q * 64 + j * 8);
          vb = *reinterpret_cast<const uint4*>(B + static_cast<size_t>(t) * 128 + q * 64 + j * 8);
        }
        *reinterpret_cast<uint4*>(sA + q * kPanelBytes + panel_chunk_offset(t, j)) = va;
        *reinterpret_cast<uint4*>(sB + q * kPanelBytes + panel_chunk_offset(t, j)) = vb;
      }
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;

  if (t == 0) {
    if (mode == 0) {
      const uint32_t idesc = umma_idesc_f16(128, 128, false, false);
      for (int p = 0; p < npan; ++p)
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ss(tmem, umma_desc_kmajor(smem_u32(sA + p * kPanelBytes), kk),
                      umma_desc_kmajor(smem_u32(sB + p * kPanelBytes), kk), idesc, (p | kk) ? 1u : 0u);
    } else if (mode == 1) {
      const uint32_t idesc = umma_idesc_f16(128, 128, true, true);
      for (int kk = 0; kk < K / 16; ++kk)
        umma_f16_ss(tmem, umma_desc_mnmajor(smem_u32(sA), kk, kPanelBytes), umma_desc_mnmajor(smem_u32(sB), kk, kPanelBytes),
                    idesc, kk ? 1u : 0u);
    } else {
      const uint32_t idesc = umma_idesc_f16(128, 16, true, true);
      const uint32_t boff = static_cast<uint32_t>(mode - 2) * 32u;
      for (int kk = 0; kk < K / 16; ++kk)
        umma_f16_ss(tmem, umma_desc_mnmajor(smem_u32(sA), kk, kPanelBytes),
                    umma_desc_mnmajor(smem_u32(sB) + boff, kk, kPanelBytes), idesc, kk ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) D[static_cast<size_t>(t) * 128 + c0 + i] = __uint_as_float(v[i]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace
}  // namespace nvp

extern "C" int nvp_selftest_umma(const void* A, const void* B, float* D, int K, int mode, void* stream) {
  using namespace nvp;
  NVP_CHECK(mode >= 0 && mode <= 5, "mode must be in [0,5]");
  if (mode == 0) NVP_CHECK(K % 64 == 0 && K >= 64 && K <= 256, "mode 0 needs K in {64,128,192,256}");
  if (mode >= 1) NVP_CHECK(K % 16 == 0 && K >= 16 && K <= 128, "mode 1 needs K % 16 == 0, K <= 128");
  const int npan = mode == 0 ? K / 64 : 2;
  const size_t smem = static_cast<size_t>(2 * npan) * tc::kPanelBytes + 1024;
  NVP_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  umma_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(A),
                                                                            static_cast<const __half*>(B), D, K, mode);
  NVP_LAUNCH_CHECK();
  return 0;
}
