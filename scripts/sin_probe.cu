// Accuracy probe of sin.approx / cos.approx (FMUL.RZ by 1/2pi + MUFU.SIN/COS) against double precision,
// per argument range, next to the Cody-Waite-reduced variant the round-1 kernels used.
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sin_probe scripts/sin_probe.cu && /tmp/sin_probe
#include <cmath>
#include <cstdio>
#include <vector>

__device__ __forceinline__ float reduce_2pi(float x) {
  const float y = x * 0.15915494309189535f;
  const float k = __fadd_rn(__fadd_rn(y, 12582912.0f), -12582912.0f);
  float r = fmaf(k, -6.2831854820251465f, x);
  return fmaf(k, 1.7484556e-7f, r);
}

__global__ void probe(const float* x, float* s0, float* c0, float* s1, float* c1, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  __sincosf(x[i], &s0[i], &c0[i]);
  __sincosf(reduce_2pi(x[i]), &s1[i], &c1[i]);
}

int main() {
  const int n = 1 << 22;
  const float ranges[] = {3.14159f, 8.f, 16.f, 32.f, 64.f, 128.f, 512.f, 4096.f};
  std::vector<float> hx(n), a(n), b(n), c(n), d(n);
  float *x, *s0, *c0, *s1, *c1;
  cudaMalloc(&x, n * 4); cudaMalloc(&s0, n * 4); cudaMalloc(&c0, n * 4); cudaMalloc(&s1, n * 4); cudaMalloc(&c1, n * 4);
  for (float R : ranges) {
    for (int i = 0; i < n; ++i) hx[i] = R * (2.0f * (static_cast<float>(i) + 0.5f) / n - 1.0f);
    cudaMemcpy(x, hx.data(), n * 4, cudaMemcpyHostToDevice);
    probe<<<n / 256, 256>>>(x, s0, c0, s1, c1, n);
    cudaMemcpy(a.data(), s0, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), c0, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(c.data(), s1, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(d.data(), c1, n * 4, cudaMemcpyDeviceToHost);
    double e[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
      const double xs = std::sin(static_cast<double>(hx[i])), xc = std::cos(static_cast<double>(hx[i]));
      e[0] = std::fmax(e[0], std::fabs(a[i] - xs)); e[1] = std::fmax(e[1], std::fabs(b[i] - xc));
      e[2] = std::fmax(e[2], std::fabs(c[i] - xs)); e[3] = std::fmax(e[3], std::fabs(d[i] - xc));
    }
    printf("|x| <= %8.2f : sin.approx %.3e cos.approx %.3e | reduced sin %.3e cos %.3e\n", R, e[0], e[1], e[2], e[3]);
  }
  return 0;
}
