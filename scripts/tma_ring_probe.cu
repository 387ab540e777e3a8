// How fast can one SM stream an L2-resident weight set through a cp.async.bulk ring?  One producer thread and one
// consumer thread per CTA, 148 CTAs; the consumer releases a stage as soon as it has landed.  Prints SM cycles per
// 16 KiB and bytes per cycle per SM for several ring shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_ring_probe scripts/tma_ring_probe.cu && /tmp/tma_ring_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// nprod producer threads (lane 0 of warps 0..nprod-1): producer q issues the loads g with g % nprod == q.
__global__ void ring(const uint8_t* src, uint32_t src_bytes, uint32_t stage_bytes, int nstage, int n_loads, long long* cycles, int nprod) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[32], empty[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < nstage; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  const int warp = threadIdx.x >> 5;
  if (warp < nprod && (threadIdx.x & 31) == 0) {
    for (int g = warp; g < n_loads; g += nprod) {
      const int st = g % nstage, ph = (g / nstage) & 1;
      const uint32_t off = (static_cast<uint32_t>(g) * stage_bytes) % src_bytes;
      mbar_wait(&empty[st], ph ^ 1);
      mbar_expect(&full[st], stage_bytes);
      bulk_g2s(smem + static_cast<size_t>(st) * stage_bytes, src + off, stage_bytes, &full[st]);
    }
  } else if (threadIdx.x == 32 * nprod) {
    for (int g = 0; g < n_loads; ++g) {
      const int st = g % nstage, ph = (g / nstage) & 1;
      mbar_wait(&full[st], ph);
      mbar_arrive(&empty[st]);
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  const uint32_t src_bytes = 448 * 1024;
  uint8_t* src; long long* cyc;
  cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaFuncSetAttribute(ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int shapes[][3] = {{16384, 1, 1}, {16384, 2, 1}, {16384, 3, 1}, {16384, 4, 1}, {16384, 6, 1}, {16384, 12, 1},
                           {8192, 3, 1}, {8192, 6, 1}, {8192, 24, 1}, {4096, 12, 1}, {32768, 2, 1}, {32768, 3, 1}, {32768, 6, 1},
                           {16384, 3, 2}, {16384, 4, 2}, {16384, 6, 2}, {16384, 6, 3}, {16384, 12, 4}, {8192, 6, 2}, {8192, 12, 4}, {4096, 12, 4},
                           {32768, 4, 2}};
  for (auto& s : shapes) {
    const uint32_t stage = s[0]; const int nst = s[1], nprod = s[2];
    const int n_loads = static_cast<int>(64ull * 1024 * 1024 / 148 / stage) + 1;   // ~64 MiB per launch over 148 SMs
    for (int rep = 0; rep < 2; ++rep) ring<<<148, 32 * nprod + 32, stage * nst>>>(src, src_bytes, stage, nst, n_loads, cyc, nprod);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (long long v : h) mx = v > mx ? v : mx;
    const double per = static_cast<double>(mx) / n_loads;
    printf("stage %6u B x %2d stages, %d producer(s) (%3u KiB in flight): %7.0f cycles per stage = %6.1f cycles per 16 KiB, %5.1f B/cycle/SM\n", stage, nst,
           nprod, stage * nst / 1024, per, per * 16384.0 / stage, stage / per);
  }
  return 0;
}
