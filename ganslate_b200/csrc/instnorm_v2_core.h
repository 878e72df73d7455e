// Per-thread body of the second-generation fused InstanceNorm backward (instnorm_v2.cu, opt-in: knob 22).
//
// The same source compiles for the device (nvcc, included by instnorm_v2.cu) and for the host (g++, included by
// tests/emul/in_bwd_v2_emul.cpp): the CPU test-suite runs every "thread" of every "block" through this code on the
// host and compares the result with torch, so the index arithmetic (row segments, reflection fold, ragged tails),
// the per-channel algebra and the bf16 packing are checked without a GPU.  Only the block reduction, the atomics and
// the grid barrier live in the .cu file.
//
// Differences from the first generation (instnorm_fast.cu, profiles/r01p_ncu_full_in_bwd_b8.md: 29 instructions per
// element and pass, 45 % issue slots busy at 33 % occupancy, long-scoreboard + barrier stalls):
//   * a thread owns 8 channels (one 16-byte bf16 vector, two 16-byte fp32 vectors) instead of 4: the per-pixel
//     addressing / predicate work is shared by twice the elements and a pixel is 48 bytes in flight per thread;
//   * the block's pixel range is walked as ROW SEGMENTS: row offsets and the row's reflection partner are computed
//     once per segment (uniform over the block), the per-pixel work is one multiply-add per view and one range test;
//   * xhat = fma(x, rstd, -mean * rstd) and dx = fma(-xhat, k2, fma(rstd, g, -k1)), k1 = rstd * mean(g),
//     k2 = rstd * mean(g * xhat): 2 + 2 arithmetic instructions per element instead of 3 + 4.
#pragma once
#include <stdint.h>
#include "../../include/ganslate_b200.h"

#if defined(__CUDACC__)
#define V2_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <string.h>
#define V2_HD inline
struct float4 { float x, y, z, w; };
struct uint4 { uint32_t x, y, z, w; };
#endif

namespace gbv2 {

constexpr int THREADS = 256;
constexpr int NO_MIRROR = -(1 << 20);

struct Geom {
  int ppb;                   // pixels per block
  int nblocks;               // blocks per image
  int total_blocks;          // grid size (single-launch mode: barrier target)
};

inline int slots_of(int C) { return THREADS / (C >> 3); }  // pixels a block covers per step

// one resident wave: `cap` co-resident blocks shared by the N images, equal pixel ranges per block
inline Geom plan(int N, int D, int H, int W, int C, int cap, bool* fits) {
  (void)C;
  Geom g;
  const int64_t P = (int64_t)D * H * W;
  int nb = cap / N;
  *fits = nb >= 1;
  if (nb < 1) nb = 1;
  int64_t ppb = (P + nb - 1) / nb;
  if (ppb < 1) ppb = 1;
  g.ppb = (int)ppb;
  g.nblocks = (int)((P + ppb - 1) / ppb);
  g.total_blocks = g.nblocks * N;
  return g;
}

// border index that reflects onto interior index i of an axis of length n with border p, or NO_MIRROR (the host only
// takes this path when n > 2p + 1: at most one mirror image per axis)
V2_HD int mirror_of(int i, int n, int p) {
  if (i >= 1 && i <= p) return -i;
  if (i <= n - 2 && i >= n - 1 - p) return 2 * (n - 1) - i;
  return NO_MIRROR;
}

V2_HD float as_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
V2_HD float rsqrt_(float v) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(v);
#else
  return 1.0f / sqrtf(v);
#endif
}
V2_HD float fma_(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
// two fp32 -> packed bf16 pair, round to nearest even
V2_HD uint32_t pack2(float lo, float hi) {
#if defined(__CUDA_ARCH__)
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
#else
  uint32_t a, b;
  memcpy(&a, &lo, 4);
  memcpy(&b, &hi, 4);
  a = (a + 0x7FFFu + ((a >> 16) & 1u)) >> 16;
  b = (b + 0x7FFFu + ((b >> 16) & 1u)) >> 16;
  return (a & 0xFFFFu) | (b << 16);
#endif
}
V2_HD void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = as_float(u.x << 16); f[1] = as_float(u.x & 0xFFFF0000u);
  f[2] = as_float(u.y << 16); f[3] = as_float(u.y & 0xFFFF0000u);
  f[4] = as_float(u.z << 16); f[5] = as_float(u.z & 0xFFFF0000u);
  f[6] = as_float(u.w << 16); f[7] = as_float(u.w & 0xFFFF0000u);
}
V2_HD float4 ld_stream4(const float* p) {  // written during this launch by other blocks (bstats): L2 only
#if defined(__CUDA_ARCH__)
  return __ldcg(reinterpret_cast<const float4*>(p));
#else
  return *reinterpret_cast<const float4*>(p);
#endif
}
// the incoming gradient: read-only for the whole launch, so the non-coherent path (L1-allocating) is safe -- and wanted:
// a thread's two 16-byte loads of a pixel are the two halves of the same 32-byte sectors, the second one hits L1
// instead of fetching the sectors from L2 again
V2_HD float4 ld_grad4(const float* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(reinterpret_cast<const float4*>(p));
#else
  return *reinterpret_cast<const float4*>(p);
#endif
}
// the residual gradient: read and rewritten by this thread only (ordinary cached load, same two-halves argument)
V2_HD float4 ld_own4(const float* p) { return *reinterpret_cast<const float4*>(p); }
V2_HD uint4 ld_bf16x8(const uint16_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(reinterpret_cast<const uint4*>(p));
#else
  return *reinterpret_cast<const uint4*>(p);
#endif
}

V2_HD void add4(float (&f)[8], int o, const float4& t) {
  f[o] += t.x; f[o + 1] += t.y; f[o + 2] += t.z; f[o + 3] += t.w;
}

// One pass of one thread over its share of block `bx` of image `n`.
//   PASS 0: acc1 / acc2 = this thread's partial (sum g, sum g * xhat) of its 8 channels; RES: dy_sum += folded gradient
//   PASS 1: dx is written; acc1 = partial sum of the fp32 dx (bias gradient); reads bstats (totals of pass 0)
// g = gradient on the (reflection-padded when dy_b.pad > 0) output domain folded onto the interior, times the
// activation's derivative (xhat > 0 ? 1 : neg_slope).
// GEN adds what the V-Net layers need (ganslate/nn/generators/vnet/vnet3d.py:160-168,193-203): per-channel PReLU
// slopes (p.prelu) and their gradient (acc3 = partial sum of g * pre over the negative side, PASS 0), a residual added
// BEFORE the activation (pre = xhat + res: the mask depends on it and dy_sum receives the MASKED gradient) and a
// scaled output (out_scale, the inverse of an additive coupling).
template <bool RES, int U, int PASS, bool GEN>
V2_HD void stream_pass(const gb_in_bwd_params& p, const Geom& g, float neg_slope, int tid, int bx, int n,
                       float (&acc1)[8], float (&acc2)[8], float (&acc3)[8]) {
  const gb_view& x = p.x;
  const gb_view& dy = p.dy_b;
  const int C8 = x.C >> 3;
  const int slots = THREADS / C8;
  const int cgp = tid % C8;
  const int slot = tid / C8;
#pragma unroll
  for (int e = 0; e < 8; ++e) acc1[e] = acc2[e] = acc3[e] = 0.f;
  if (slot >= slots) return;
  const int c = cgp * 8;
  const int W = x.W;
  const uint32_t P = (uint32_t)x.D * (uint32_t)x.H * (uint32_t)x.W;
  const uint32_t p0 = (uint32_t)bx * (uint32_t)g.ppb;
  const uint32_t p1 = (p0 + (uint32_t)g.ppb < P) ? p0 + (uint32_t)g.ppb : P;
  if (p0 >= p1) return;
  const float invP = 1.f / (float)P;

  float a[8], b[8], k1[8], k2[8];
  {
    const float* sp = p.stats + ((int64_t)n * x.C + c) * 2;
    const float* bp = p.bstats + ((int64_t)n * x.C + c) * 2;
#pragma unroll
    for (int h = 0; h < 4; ++h) {  // (sum, sum of squares) of channels 2h, 2h + 1
      const float4 s = *reinterpret_cast<const float4*>(sp + 4 * h);
      const float m0 = s.x * invP, m1 = s.z * invP;
      const float v0 = s.y * invP - m0 * m0, v1 = s.w * invP - m1 * m1;
      const float r0 = rsqrt_((v0 > 0.f ? v0 : 0.f) + p.eps), r1 = rsqrt_((v1 > 0.f ? v1 : 0.f) + p.eps);
      a[2 * h] = r0;
      a[2 * h + 1] = r1;
      b[2 * h] = -m0 * r0;
      b[2 * h + 1] = -m1 * r1;
      if (PASS == 1) {
        const float4 t = ld_stream4(bp + 4 * h);  // (sum g, sum g * xhat) written by pass 0 of every block
        k1[2 * h] = r0 * (t.x * invP);
        k2[2 * h] = r0 * (t.y * invP);
        k1[2 * h + 1] = r1 * (t.z * invP);
        k2[2 * h + 1] = r1 * (t.w * invP);
      } else {
        k1[2 * h] = k2[2 * h] = k1[2 * h + 1] = k2[2 * h + 1] = 0.f;
      }
    }
  }
  const bool want_dbias = PASS == 1 && p.dbias != nullptr;
  const bool do_res = RES && PASS == 0;  // the residual gradient is accumulated exactly once
  const bool rba = GEN && p.res_before_act != 0 && p.res.ptr != nullptr;
  const bool want_dprelu = GEN && PASS == 0 && p.dprelu != nullptr;
  const float oscale = (GEN && p.out_scale != 0.f) ? p.out_scale : 1.f;
  float ns[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) ns[e] = (GEN && p.prelu != nullptr) ? p.prelu[c + e] : neg_slope;
  const uint16_t* rb = rba ? reinterpret_cast<const uint16_t*>(p.res.ptr) + (int64_t)n * p.res.sn + c : nullptr;
  const int rsx = rba ? (int)p.res.sx : 0;
  const float* gb = reinterpret_cast<const float*>(dy.ptr) + (int64_t)n * dy.sn + c;
  const uint16_t* xb = reinterpret_cast<const uint16_t*>(x.ptr) + (int64_t)n * x.sn + c;
  float* sb = RES ? reinterpret_cast<float*>(p.dy_sum.ptr) + (int64_t)n * p.dy_sum.sn + c : nullptr;
  uint16_t* db = reinterpret_cast<uint16_t*>(p.dx.ptr) + (int64_t)n * p.dx.sn + c;
  const int gpad = dy.pad;
  const int gsx = (int)dy.sx, gsy = (int)dy.sy, xsx = (int)x.sx, dsx = (int)p.dx.sx;
  const int ssx = RES ? (int)p.dy_sum.sx : 0;

  // Row segments [xa, xe) of the block's pixel range, one per trip (uniform over the block).  The apply pass walks
  // them in REVERSE: when the two passes run in one launch it starts with the rows the reduction pass read last --
  // the part of the tensors most likely still in L2 (an LRU cache re-read in the same order would miss everywhere
  // once the tensors exceed it: 268 MB of gradient + x on the 64-channel 256 x 256 layers at batch 8).
  const int y_first = (int)(p0 / (uint32_t)W), y_last = (int)((p1 - 1u) / (uint32_t)W);
  const int x_first = (int)(p0 - (uint32_t)y_first * (uint32_t)W);
  const int x_end_last = (int)(p1 - (uint32_t)y_last * (uint32_t)W);  // exclusive end in the last row
  for (int r = 0; r <= y_last - y_first; ++r) {
    const int y = PASS == 1 ? y_last - r : y_first + r;
    const int xa = y == y_first ? x_first : 0;
    const int xe = y == y_last ? x_end_last : W;
    const int my = gpad > 0 ? mirror_of(y, dy.H, gpad) : NO_MIRROR;
    const int og = y * gsy, ox = y * (int)x.sy, od = y * (int)p.dx.sy;
    const int os = RES ? y * (int)p.dy_sum.sy : 0;
    const int orr = rba ? y * (int)p.res.sy : 0;
    for (int px = xa + slot; px < xe; px += slots * U) {
      float4 g0[U], g1[U], q0[U], q1[U];
      uint4 xv[U], rv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pxu = px + u * slots;
        if (pxu < xe) {
          const float* gp = gb + og + pxu * gsx;
          g0[u] = ld_grad4(gp);
          g1[u] = ld_grad4(gp + 4);
          xv[u] = ld_bf16x8(xb + ox + pxu * xsx);
          if (rba) rv[u] = ld_bf16x8(rb + orr + pxu * rsx);
          if (do_res) {
            const float* rp = sb + os + pxu * ssx;
            q0[u] = ld_own4(rp);
            q1[u] = ld_own4(rp + 4);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pxu = px + u * slots;
        if (pxu < xe) {
          float gg[8] = {g0[u].x, g0[u].y, g0[u].z, g0[u].w, g1[u].x, g1[u].y, g1[u].z, g1[u].w};
          float xf[8];
          unpack8(xv[u], xf);
          if (gpad > 0) {
            // columns 1..gpad and W-1-gpad..W-2 have a mirror image in the border; rows likewise (my)
            const bool col_edge = (unsigned)(pxu - 1) < (unsigned)gpad || (unsigned)(W - 2 - pxu) < (unsigned)gpad;
            if (my != NO_MIRROR || col_edge) {
              const int mx = col_edge ? mirror_of(pxu, W, gpad) : NO_MIRROR;
              if (my != NO_MIRROR) {
                const float* q = gb + my * gsy + pxu * gsx;
                add4(gg, 0, ld_grad4(q));
                add4(gg, 4, ld_grad4(q + 4));
              }
              if (mx != NO_MIRROR) {
                const float* q = gb + og + mx * gsx;
                add4(gg, 0, ld_grad4(q));
                add4(gg, 4, ld_grad4(q + 4));
                if (my != NO_MIRROR) {
                  const float* qc = gb + my * gsy + mx * gsx;
                  add4(gg, 0, ld_grad4(qc));
                  add4(gg, 4, ld_grad4(qc + 4));
                }
              }
            }
          }
          if (do_res && !rba) {  // residual added after the activation: it receives the unmasked gradient
            float* rp = sb + os + pxu * ssx;
            float4 o0, o1;
            o0.x = q0[u].x + gg[0]; o0.y = q0[u].y + gg[1]; o0.z = q0[u].z + gg[2]; o0.w = q0[u].w + gg[3];
            o1.x = q1[u].x + gg[4]; o1.y = q1[u].y + gg[5]; o1.z = q1[u].z + gg[6]; o1.w = q1[u].w + gg[7];
            *reinterpret_cast<float4*>(rp) = o0;
            *reinterpret_cast<float4*>(rp + 4) = o1;
          }
          float d[8], rf[8];
          if (rba) unpack8(rv[u], rf);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float xh = fma_(xf[e], a[e], b[e]);
            float ge = gg[e];
            if (GEN) {
              const float pre = rba ? xh + rf[e] : xh;
              ge *= oscale;
              if (!(pre > 0.f)) {
                if (want_dprelu) acc3[e] = fma_(ge, pre, acc3[e]);
                ge *= ns[e];
              }
              gg[e] = ge;  // masked gradient (residual before the activation)
            } else {
              if (!(xh > 0.f)) ge *= neg_slope;
            }
            if (PASS == 0) {
              acc1[e] += ge;
              acc2[e] = fma_(ge, xh, acc2[e]);
            } else {
              d[e] = fma_(-xh, k2[e], fma_(a[e], ge, -k1[e]));
              if (want_dbias) acc1[e] += d[e];  // sum of the fp32 dx, not of its bf16 rounding (see instnorm.cu)
            }
          }
          if (do_res && rba) {
            float* rp = sb + os + pxu * ssx;
            float4 o0, o1;
            o0.x = q0[u].x + gg[0]; o0.y = q0[u].y + gg[1]; o0.z = q0[u].z + gg[2]; o0.w = q0[u].w + gg[3];
            o1.x = q1[u].x + gg[4]; o1.y = q1[u].y + gg[5]; o1.z = q1[u].z + gg[6]; o1.w = q1[u].w + gg[7];
            *reinterpret_cast<float4*>(rp) = o0;
            *reinterpret_cast<float4*>(rp + 4) = o1;
          }
          if (PASS == 1) {
            uint4 o;
            o.x = pack2(d[0], d[1]);
            o.y = pack2(d[2], d[3]);
            o.z = pack2(d[4], d[5]);
            o.w = pack2(d[6], d[7]);
            *reinterpret_cast<uint4*>(db + od + pxu * dsx) = o;
          }
        }
      }
    }
  }
}

}  // namespace gbv2

// ------------------------------------------------------------------------------------------------ forward
// y = [out_scale *] act(xhat [+ res]) [+ res] incl. the reflection border of y (y.pad > 0, 2-D): same walk as the
// backward (row segments, 8 channels per thread, U pixels in flight).  GEN: PReLU slopes, residual before the
// activation, scaled output -- the V-Net forms, which the first-generation fast kernel leaves to the general one.
namespace gbv2 {

V2_HD void st_bf16x8(uint16_t* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

template <int U, bool GEN>
V2_HD void fwd_pass(const gb_in_fwd_params& p, const Geom& g, float neg_slope, int tid, int bx, int n) {
  const gb_view& x = p.x;
  const gb_view& yv = p.y;
  const int C8 = x.C >> 3;
  const int slots = THREADS / C8;
  const int cgp = tid % C8;
  const int slot = tid / C8;
  if (slot >= slots) return;
  const int c = cgp * 8;
  const int W = x.W;
  const uint32_t P = (uint32_t)x.D * (uint32_t)x.H * (uint32_t)x.W;
  const uint32_t p0 = (uint32_t)bx * (uint32_t)g.ppb;
  const uint32_t p1 = (p0 + (uint32_t)g.ppb < P) ? p0 + (uint32_t)g.ppb : P;
  if (p0 >= p1) return;
  float a[8], b[8], ns[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    a[e] = 1.f;
    b[e] = 0.f;
    ns[e] = (GEN && p.prelu != nullptr) ? p.prelu[c + e] : neg_slope;
  }
  if (p.stats != nullptr) {
    const float invP = 1.f / (float)P;
    const float* sp = p.stats + ((int64_t)n * x.C + c) * 2;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float4 s = *reinterpret_cast<const float4*>(sp + 4 * h);
      const float m0 = s.x * invP, m1 = s.z * invP;
      const float v0 = s.y * invP - m0 * m0, v1 = s.w * invP - m1 * m1;
      const float r0 = rsqrt_((v0 > 0.f ? v0 : 0.f) + p.eps), r1 = rsqrt_((v1 > 0.f ? v1 : 0.f) + p.eps);
      a[2 * h] = r0;
      a[2 * h + 1] = r1;
      b[2 * h] = -m0 * r0;
      b[2 * h + 1] = -m1 * r1;
    }
  }
  const bool has_res = p.res.ptr != nullptr;
  const bool rba = GEN && has_res && p.res_before_act != 0;
  const float oscale = (GEN && p.out_scale != 0.f) ? p.out_scale : 1.f;
  const uint16_t* xb = reinterpret_cast<const uint16_t*>(x.ptr) + (int64_t)n * x.sn + c;
  const uint16_t* rb = has_res ? reinterpret_cast<const uint16_t*>(p.res.ptr) + (int64_t)n * p.res.sn + c : nullptr;
  uint16_t* yb = reinterpret_cast<uint16_t*>(yv.ptr) + (int64_t)n * yv.sn + c;
  const int ypad = yv.pad;
  const int xsx = (int)x.sx, ysx = (int)yv.sx, ysy = (int)yv.sy, rsx = has_res ? (int)p.res.sx : 0;

  int y = (int)(p0 / (uint32_t)W);
  int xa = (int)(p0 - (uint32_t)y * (uint32_t)W);
  uint32_t pix = p0;
  while (pix < p1) {
    const int left = (int)(p1 - pix);
    const int seg = left < W - xa ? left : W - xa;
    const int xe = xa + seg;
    const int my = ypad > 0 ? mirror_of(y, yv.H, ypad) : NO_MIRROR;
    const int ox = y * (int)x.sy, oy = y * ysy, orr = has_res ? y * (int)p.res.sy : 0;
    for (int px = xa + slot; px < xe; px += slots * U) {
      uint4 xv[U], rv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pxu = px + u * slots;
        if (pxu < xe) {
          xv[u] = ld_bf16x8(xb + ox + pxu * xsx);
          if (has_res) rv[u] = ld_bf16x8(rb + orr + pxu * rsx);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pxu = px + u * slots;
        if (pxu < xe) {
          float f[8], r[8];
          unpack8(xv[u], f);
          if (has_res) unpack8(rv[u], r);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float v = fma_(f[e], a[e], b[e]);
            if (rba) v += r[e];
            v = v > 0.f ? v : v * ns[e];
            if (GEN) v *= oscale;
            if (has_res && !rba) v += r[e];
            f[e] = v;
          }
          uint4 o;
          o.x = pack2(f[0], f[1]);
          o.y = pack2(f[2], f[3]);
          o.z = pack2(f[4], f[5]);
          o.w = pack2(f[6], f[7]);
          st_bf16x8(yb + oy + pxu * ysx, o);
          if (ypad > 0) {
            const bool col_edge = (unsigned)(pxu - 1) < (unsigned)ypad || (unsigned)(W - 2 - pxu) < (unsigned)ypad;
            if (my != NO_MIRROR || col_edge) {
              const int mx = col_edge ? mirror_of(pxu, W, ypad) : NO_MIRROR;
              if (my != NO_MIRROR) st_bf16x8(yb + my * ysy + pxu * ysx, o);
              if (mx != NO_MIRROR) {
                st_bf16x8(yb + oy + mx * ysx, o);
                if (my != NO_MIRROR) st_bf16x8(yb + my * ysy + mx * ysx, o);
              }
            }
          }
        }
      }
    }
    pix += (uint32_t)seg;
    ++y;
    xa = 0;
  }
}

}  // namespace gbv2
