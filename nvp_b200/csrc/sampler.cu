// Device-resident coordinate sampler (reference: dataio.py:85-99,104-120 and the per-step
// host gather + H2D of training.py:45-48).  The uint8 video lives in HBM; one kernel produces a batch:
//   t_idx ~ U{t_lo..t_hi-1}, p_idx ~ U{0..H*W-1} (with replacement)
//   all_coords = [temporal_coords[t_idx], row/(H-1), col/(W-1)], temporal_steps = temporal_steps[t_idx],
//   img = video[t_idx, p_idx, :]
// Two index sources: explicit index arrays (the reference's CPU mt19937 stream, uploaded: "parity mode") or a
// counter-based Philox4x32-10 generator keyed by (seed, step, sample) ("throughput mode", a different stream).
#include "common.cuh"

namespace nvp {
namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c[0];
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c[2];
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c[3] ^ k1;
    c[1] = static_cast<uint32_t>(p1); c[3] = static_cast<uint32_t>(p0);
    c[0] = n0; c[2] = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

struct SampleArgs {
  const uint8_t* video;      // [T, H*W, 3]
  const float* tcoords;      // [T] = linspace(0, 1, T)               dataio.py:99
  const float* tsteps_lut;   // [T] = linspace(.5/T, 1-.5/T, T)       dataio.py:96
  const int64_t* t_idx_in;   // optional explicit indices
  const int64_t* p_idx_in;
  int T, H, W, t_lo, t_hi;
  int64_t n;
  uint64_t seed, step;
  float* coords; float* tsteps; uint8_t* gt;
  int32_t* t_idx_out; int32_t* p_idx_out;   // optional
};

__global__ void __launch_bounds__(256) sample_kernel(const SampleArgs a) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  int t, p;
  const int hw = a.H * a.W;
  if (a.t_idx_in != nullptr) {
    t = static_cast<int>(a.t_idx_in[i]);
    p = static_cast<int>(a.p_idx_in[i]);
  } else {
    uint32_t c[4] = {static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), static_cast<uint32_t>(a.step),
                     static_cast<uint32_t>(a.step >> 32)};
    philox4x32_10(c, static_cast<uint32_t>(a.seed), static_cast<uint32_t>(a.seed >> 32));
    t = a.t_lo + static_cast<int>((static_cast<uint64_t>(c[0]) * static_cast<uint32_t>(a.t_hi - a.t_lo)) >> 32);
    p = static_cast<int>((static_cast<uint64_t>(c[1]) * static_cast<uint32_t>(hw)) >> 32);
  }
  const int row = p / a.W, col = p - row * a.W;
  a.coords[3 * i] = __ldg(a.tcoords + t);
  a.coords[3 * i + 1] = __fdiv_rn(static_cast<float>(row), static_cast<float>(a.H - 1));   // get_mgrid, dataio.py:17-19
  a.coords[3 * i + 2] = __fdiv_rn(static_cast<float>(col), static_cast<float>(a.W - 1));
  a.tsteps[i] = __ldg(a.tsteps_lut + t);
  const uint8_t* px = a.video + (static_cast<size_t>(t) * hw + p) * 3;
  a.gt[3 * i] = __ldg(px); a.gt[3 * i + 1] = __ldg(px + 1); a.gt[3 * i + 2] = __ldg(px + 2);
  if (a.t_idx_out) { a.t_idx_out[i] = t; a.p_idx_out[i] = p; }
}

}  // namespace
}  // namespace nvp

extern "C" int nvp_sample_batch(const uint8_t* video, int T, int H, int W, const float* temporal_coords,
                                const float* temporal_steps, int64_t n, const int64_t* t_idx, const int64_t* p_idx,
                                uint64_t seed, uint64_t step, int t_lo, int t_hi, float* coords, float* tsteps, uint8_t* gt,
                                int32_t* t_idx_out, int32_t* p_idx_out, void* stream) {
  using namespace nvp;
  reset_launch_count();
  NVP_CHECK(video && temporal_coords && temporal_steps && coords && tsteps && gt, "nvp_sample_batch: NULL buffer");
  NVP_CHECK(T >= 1 && H >= 2 && W >= 2 && static_cast<int64_t>(H) * W < (1ll << 31), "nvp_sample_batch: bad video shape");
  NVP_CHECK((t_idx == nullptr) == (p_idx == nullptr), "nvp_sample_batch: give both index arrays or neither");
  NVP_CHECK(0 <= t_lo && t_lo < t_hi && t_hi <= T, "nvp_sample_batch: need 0 <= t_lo < t_hi <= T");
  NVP_CHECK((t_idx_out == nullptr) == (p_idx_out == nullptr), "nvp_sample_batch: give both index outputs or neither");
  if (n <= 0) return 0;
  SampleArgs a{};
  a.video = video; a.tcoords = temporal_coords; a.tsteps_lut = temporal_steps; a.t_idx_in = t_idx; a.p_idx_in = p_idx;
  a.T = T; a.H = H; a.W = W; a.t_lo = t_lo; a.t_hi = t_hi; a.n = n; a.seed = seed; a.step = step;
  a.coords = coords; a.tsteps = tsteps; a.gt = gt; a.t_idx_out = t_idx_out; a.p_idx_out = p_idx_out;
  ScopedKernelTimer timer(K_MISC, static_cast<cudaStream_t>(stream));
  sample_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  NVP_LAUNCH_CHECK();
  return 0;
}
