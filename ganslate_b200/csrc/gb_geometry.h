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

struct gb_row {
  int n, qz, qy, qx;
};

GB_HD gb_row gb_decode_row(int64_t m, const int (&q)[3]) {
  gb_row r;
  r.qx = (int)(m % q[2]);
  m /= q[2];
  r.qy = (int)(m % q[1]);
  m /= q[1];
  r.qz = (int)(m % q[0]);
  r.n = (int)(m / q[0]);
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
