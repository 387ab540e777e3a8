// fp16 latent-row access helpers of the grid kernels (included by grid.cu inside nvp::<anonymous>).
#pragma once

// F consecutive halfs (one level of one plane) of a latent row in the MMA tile format; never straddles a 16-byte chunk
// because F divides 8 and the column is a multiple of F.  The load returns the raw bits (so that a prefetch does not
// wait for the data); cvt_halfs converts at the point of use.
template <int F> struct RawHalfs { uint32_t w[(F + 1) / 2]; };
template <int F>
__device__ __forceinline__ RawHalfs<F> ld_halfs_raw(const uint8_t* p) {
  RawHalfs<F> q;
  if constexpr (F == 1) {
    q.w[0] = __ldg(reinterpret_cast<const unsigned short*>(p));
  } else if constexpr (F == 2) {
    q.w[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
  } else if constexpr (F == 4) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    q.w[0] = t.x; q.w[1] = t.y;
  } else {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    q.w[0] = t.x; q.w[1] = t.y; q.w[2] = t.z; q.w[3] = t.w;
  }
  return q;
}
// the same from a shared-memory staging buffer
template <int F>
__device__ __forceinline__ RawHalfs<F> lds_halfs_raw(const uint8_t* p) {
  RawHalfs<F> q;
  if constexpr (F == 1) {
    q.w[0] = *reinterpret_cast<const unsigned short*>(p);
  } else if constexpr (F == 2) {
    q.w[0] = *reinterpret_cast<const uint32_t*>(p);
  } else if constexpr (F == 4) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    q.w[0] = t.x; q.w[1] = t.y;
  } else {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    q.w[0] = t.x; q.w[1] = t.y; q.w[2] = t.z; q.w[3] = t.w;
  }
  return q;
}
template <int F>
__device__ __forceinline__ void cvt_halfs(const RawHalfs<F>& q, float (&v)[F]) {
  if constexpr (F == 1) {
    v[0] = __half2float(__ushort_as_half(static_cast<unsigned short>(q.w[0])));
  } else {
#pragma unroll
    for (int i = 0; i < F / 2; ++i) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&q.w[i]));
      v[2 * i] = t.x; v[2 * i + 1] = t.y;
    }
  }
}
template <int F>
__device__ __forceinline__ void st_halfs(uint8_t* p, const float (&v)[F]) {
  if constexpr (F == 1) {
    *reinterpret_cast<__half*>(p) = __float2half_rn(v[0]);
  } else if constexpr (F == 2) {
    *reinterpret_cast<uint32_t*>(p) = tc::pack_half2(v[0], v[1]);
  } else if constexpr (F == 4) {
    *reinterpret_cast<uint2*>(p) = make_uint2(tc::pack_half2(v[0], v[1]), tc::pack_half2(v[2], v[3]));
  } else {
    *reinterpret_cast<uint4*>(p) = make_uint4(tc::pack_half2(v[0], v[1]), tc::pack_half2(v[2], v[3]),
                                              tc::pack_half2(v[4], v[5]), tc::pack_half2(v[6], v[7]));
  }
}

