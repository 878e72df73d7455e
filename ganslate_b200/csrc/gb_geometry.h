// Geometry of the implicit GEMMs, shared by the CUDA kernels and by the host-side emulator that the
// CPU tests use to check the index math (tests/ only -- never on the product path).
#pragma once
#include <stdint.h>
#include "../../include/ganslate_b200.h"

#if defined(__CUDACC__)
#define GB_HD __host__ __device__ __forceinline__
#else
#define GB_HD inline
#endif

// extents of the q-grid of one class: q ranges over ceil((out_extent - off) / out_mul)
GB_HD void gb_class_extents(const gb_conv_params& p, int cls, int (&q)[3]) {
  const int ext[3] = {p.out.D, p.out.H, p.out.W};
  for (int i = 0; i < 3; ++i) {
    int e = ext[i] - p.cls[cls].off[i];
    q[i] = e <= 0 ? 0 : (e + p.out_mul[i] - 1) / p.out_mul[i];
  }
}

// Division by a launch-invariant divisor without an integer divide (Granlund-Montgomery round-up method):
//   x / d == (umulhi(x, mul) + x) >> shr   for 0 <= x < 2^31, 1 <= d < 2^31.
struct gb_fastdiv {
  uint32_t mul, shr, d, pad_;
};
inline gb_fastdiv gb_make_fastdiv(uint32_t d) {
  gb_fastdiv f;
  f.d = d ? d : 1;
  uint32_t l = 0;
  while ((1ull << l) < f.d) ++l;
  f.mul = (uint32_t)(((1ull << 32) * ((1ull << l) - f.d)) / f.d + 1);
  f.shr = l;
  f.pad_ = 0;
  return f;
}
// one definition for device code and for host code / CPU tests that replay the kernels' tile decoding
GB_HD uint32_t gb_div(uint32_t x, const gb_fastdiv& f) {
#if defined(__CUDA_ARCH__)
  return (__umulhi(x, f.mul) + x) >> f.shr;
#else
  return ((uint32_t)(((uint64_t)x * f.mul) >> 32) + x) >> f.shr;
#endif
}

struct gb_row {
  int n, qz, qy, qx;
};

#if defined(__CUDACC__)
// decode with precomputed magic numbers: f[0..2] divide by the (z, y, x) extents of the grid
__device__ __forceinline__ gb_row gb_decode_row_fast(uint32_t m, const gb_fastdiv* f) {
  gb_row r;
  uint32_t t = gb_div(m, f[2]);
  r.qx = (int)(m - t * f[2].d);
  m = t;
  t = gb_div(m, f[1]);
  r.qy = (int)(m - t * f[1].d);
  m = t;
  t = gb_div(m, f[0]);
  r.qz = (int)(m - t * f[0].d);
  r.n = (int)t;
  return r;
}
#endif

// 32-bit arithmetic on purpose: row indices are checked < 2^31 on the host and 64-bit integer division costs
// ~100 instructions on the GPU (it dominated the first version of the wgrad producer loop).
GB_HD gb_row gb_decode_row(int64_t m64, const int (&q)[3]) {
  gb_row r;
  uint32_t m = (uint32_t)m64;
  const uint32_t q2 = (uint32_t)q[2], q1 = (uint32_t)q[1], q0 = (uint32_t)q[0];
  uint32_t t = m / q2;
  r.qx = (int)(m - t * q2);
  m = t;
  t = m / q1;
  r.qy = (int)(m - t * q1);
  m = t;
  t = m / q0;
  r.qz = (int)(m - t * q0);
  r.n = (int)t;
  return r;
}

// element offset of pixel (n,z,y,x) channel 0 inside a view; caller checks bounds
GB_HD int64_t gb_pix_offset(const gb_view& v, int n, int z, int y, int x) {
  return (int64_t)n * v.sn + (int64_t)z * v.sz + (int64_t)y * v.sy + (int64_t)x * v.sx;
}

GB_HD bool gb_in_bounds(const gb_view& v, int z, int y, int x) {
  return (unsigned)z < (unsigned)v.D && (unsigned)y < (unsigned)v.H && (unsigned)x < (unsigned)v.W;
}

// reflect index i into [0, n) (PyTorch ReflectionPad semantics, single bounce)
GB_HD int gb_reflect(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
