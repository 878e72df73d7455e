// Element bodies of the second-generation weight pack / weight-gradient unpack kernels (pack_v2.cu, opt-in: knob 28).
// The first generation (pack.cu) walks a 64-bit flat index with three (pack) / four (unpack) 64-bit divisions per
// ELEMENT and 2-byte stores: the launches are instruction-bound (pack_multi 77 us for 91 MB of traffic, unpack_multi
// 28 us; profiles/r01p_launches_b8.md).  Here all index arithmetic is 32-bit, and the pack produces eight
// consecutive K positions (same tap, consecutive channels: chans_pad % 8 == 0) per thread and stores them as one
// 16-byte vector.  Same source for nvcc and g++ (tests/emul/pack_v2_emul.cpp), see instnorm_v2_core.h.
#pragma once
#include "instnorm_v2_core.h"

namespace gbp2 {

// eight packed bf16 of class `cls`: row n = i8 / (kpad / 8), K positions (i8 % (kpad / 8)) * 8 .. + 7
V2_HD uint4 pack8(const gb_pack_params& p, int cls, uint32_t i8) {
  const uint32_t k8n = (uint32_t)p.kpad[cls] >> 3;
  const uint32_t n = i8 / k8n;
  const uint32_t k = (i8 - n * k8n) << 3;
  const uint32_t tl = k / (uint32_t)p.chans_pad;
  const uint32_t c0 = k - tl * (uint32_t)p.chans_pad;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if ((int)n < p.rows && (int)tl < p.ntaps[cls]) {
    const int t = p.tap_id[p.tap_begin[cls] + (int)tl];  // < 0: a padding tap of a pixel-window layout (stays zero)
    if (t >= 0) {
      const float* s = p.src + (int64_t)n * p.sn + (int64_t)t * p.st + (int64_t)c0 * p.sc;
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if ((int)c0 + e < p.chans) v[e] = s[(int64_t)e * p.sc];
    }
  }
  uint4 o;
  o.x = gbv2::pack2(v[0], v[1]);
  o.y = gbv2::pack2(v[2], v[3]);
  o.z = gbv2::pack2(v[4], v[5]);
  o.w = gbv2::pack2(v[6], v[7]);
  return o;
}

// element i (destination order: row, channel, tap) of one unpack item
V2_HD void unpack1(const gb_unpack_item& it, uint32_t i) {
  const uint32_t nt = (uint32_t)it.ntaps, nc = (uint32_t)it.chans;
  const uint32_t rc = i / nt, t = i - rc * nt;
  const uint32_t r = rc / nc, c = rc - r * nc;
  const float v = it.dw[(int64_t)r * it.kpad + (int64_t)t * it.chans_pad + c];
  float* d = it.dst + (int64_t)r * it.dsr + (int64_t)c * it.dsc + (int64_t)t * it.dst_t;
  *d = it.accumulate ? *d + v : v;
}

}  // namespace gbp2
