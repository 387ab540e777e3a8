// Fused dense AdamW step over flat fp32 buffers (reference: training.py:13-14,73-76 =
// torch.optim.AdamW(lr, weight_decay=1e-3, betas=(0.9,0.999), eps=1e-8) driven by CosineAnnealingLR).
// One pass: p, g, m, v are read once, p, m, v written once and g optionally zeroed (the reference's separate
// optim.zero_grad()): 8 x 4 B per parameter instead of ~10 elementwise passes of torch 1.11's unfused AdamW.
// Arithmetic follows torch's single-tensor AdamW in the same operation order:
//   p *= 1 - lr*wd ; m = lerp(m, g, 1-b1) ; v = b2*v + (1-b2)*g*g ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace nvp {
namespace {

struct AdamArgs {
  float* p; float* g; float* m; float* v;
  int64_t n;
  float lr, beta1, beta2, eps, wd, step_size, bc2_sqrt;
  int zero_grad;
};

__device__ __forceinline__ void adamw_one(float& p, float& g, float& m, float& v, const AdamArgs& a) {
  p = p * (1.0f - a.lr * a.wd);
  m = m + (g - m) * (1.0f - a.beta1);
  v = v * a.beta2 + (1.0f - a.beta2) * g * g;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = p - a.step_size * (m / denom);
  if (a.zero_grad) g = 0.0f;
}

__global__ void __launch_bounds__(256) adamw_kernel(const AdamArgs a) {
  const int64_t n4 = a.n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(a.p)[i], g = reinterpret_cast<float4*>(a.g)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i], v = reinterpret_cast<float4*>(a.v)[i];
    adamw_one(p.x, g.x, m.x, v.x, a); adamw_one(p.y, g.y, m.y, v.y, a);
    adamw_one(p.z, g.z, m.z, v.z, a); adamw_one(p.w, g.w, m.w, v.w, a);
    reinterpret_cast<float4*>(a.p)[i] = p; reinterpret_cast<float4*>(a.m)[i] = m; reinterpret_cast<float4*>(a.v)[i] = v;
    if (a.zero_grad) reinterpret_cast<float4*>(a.g)[i] = g;
  }
  for (int64_t i = (n4 << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += stride)
    adamw_one(a.p[i], a.g[i], a.m[i], a.v[i], a);
}

}  // namespace
}  // namespace nvp

extern "C" int nvp_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int64_t step, int zero_grad,
                              void* stream) {
  using namespace nvp;
  reset_launch_count();
  NVP_CHECK(n >= 0 && step >= 1, "need n >= 0 and step >= 1");
  if (n == 0) return 0;
  NVP_CHECK(params && grads && exp_avg && exp_avg_sq, "nvp_adamw_step: NULL buffer");
  NVP_CHECK((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
             reinterpret_cast<uintptr_t>(exp_avg_sq)) % 16 == 0, "nvp_adamw_step: buffers must be 16-byte aligned");
  AdamArgs a{};
  a.p = params; a.g = grads; a.m = exp_avg; a.v = exp_avg_sq; a.n = n;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay; a.zero_grad = zero_grad;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  a.step_size = static_cast<float>(static_cast<double>(lr) / bc1);
  a.bc2_sqrt = static_cast<float>(sqrt(bc2));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = static_cast<int>(std::min<int64_t>((n / 4 + 255) / 256 + 1, static_cast<int64_t>(sms) * 16));
  ScopedKernelTimer timer(K_MISC, static_cast<cudaStream_t>(stream));
  adamw_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  NVP_LAUNCH_CHECK();
  return 0;
}
