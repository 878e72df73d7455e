// Per-thread body of the ON-CHIP fused InstanceNorm backward (instnorm_v3.cu, opt-in: knob 24).
//
// The two-pass kernels (instnorm_fast.cu, instnorm_v2.cu) read the fp32 gradient and x twice -- 14 B moved per 8 B
// of algorithmic traffic -- because a block owns a pixel range of ALL channels and the per-(image, channel) sums need
// every block of the image.  Statistics are per channel, so the work is re-partitioned here: a thread-block CLUSTER
// of K CTAs owns (image n, 32 consecutive channels) over ALL pixels, CTA r of the cluster the r-th pixel range.  Each
// CTA streams its <= 1024 pixels x 32 channels ONCE from global memory (64 B of bf16 x + 128 B of fp32 gradient per
// pixel: whole sectors), keeps the masked gradient (fp32) and x (bf16) in shared memory (192 KB), the K CTAs exchange
// their partial (sum g, sum g * xhat) through distributed shared memory, and dx is produced from the shared-memory
// copy: every byte is read once and written once, there is no grid barrier and no cooperative launch.  Covers
// feature maps of up to 8192 pixels (the 64 x 64 residual-block layers -- 18 of the 23 InstanceNorm layers of
// Resnet2D -- and the PatchGAN layers); larger maps exceed the on-chip capacity of a cluster and keep the two-pass
// kernels.
// K = smallest power of two that fits the map on chip, doubled while the launch would leave SMs idle.
//
// Same source for nvcc and for g++ (tests/emul/in_bwd_v3_emul.cpp): see instnorm_v2_core.h.
#pragma once
#include "instnorm_v2_core.h"
#include "gb_geometry.h"

namespace gbv3 {

constexpr int THREADS = 256;
constexpr int CG = 32;                      // channels per cluster
constexpr int TPP = CG / 8;                 // threads per pixel (8 channels each)
constexpr int SLOTS = THREADS / TPP;        // pixels per step of a CTA
constexpr int MAX_STEPS = 16;               // steps a thread can stash: 16 x 48 B x 256 threads = 192 KB
constexpr int MAX_PPC = SLOTS * MAX_STEPS;  // pixels per CTA
constexpr int MAX_K = 8;                    // portable cluster size

struct Geom {
  gb_fastdiv divw;           // pixel index -> row
  int K;                     // CTAs per cluster = pixel ranges per image
  int ppc;                   // pixels per CTA
  int steps;                 // ceil(ppc / SLOTS) <= MAX_STEPS
  int pad_;
};

// K: enough CTAs to hold the image on chip, then more (up to 8) while the launch would not fill `sms` SMs
// mode 2: one doubling more where possible, so that a CTA's stash is <= 96 KB and two CTAs share an SM (the load
// phase of one overlaps the exchange / apply phase of the other)
inline bool plan(int N, int D, int H, int W, int C, int sms, int mode, Geom* g) {
  const int64_t P = (int64_t)D * H * W;
  if (C % CG != 0 || P < 1 || P > (int64_t)MAX_K * MAX_PPC) return false;
  int K = 1;
  while ((int64_t)K * MAX_PPC < P) K *= 2;
  if (mode == 2 && K < MAX_K && (P + 2 * K - 1) / (2 * K) >= 2 * SLOTS) K *= 2;
  while (K < MAX_K && (int64_t)N * (C / CG) * K < sms && (P + 2 * K - 1) / (2 * K) >= 2 * SLOTS) K *= 2;
  g->K = K;
  g->ppc = (int)((P + K - 1) / K);
  g->steps = (g->ppc + SLOTS - 1) / SLOTS;
  g->divw = gb_make_fastdiv((uint32_t)W);
  g->pad_ = 0;
  return g->steps <= MAX_STEPS;
}

struct Consts {
  float a[8], b[8];          // xhat = fma(x, a, b): a = rstd, b = -mean * rstd
};

V2_HD void load_consts(const gb_in_bwd_params& p, int n, int c, Consts& k) {
  const gb_view& x = p.x;
  const float invP = 1.f / (float)((uint32_t)x.D * (uint32_t)x.H * (uint32_t)x.W);
  const float* sp = p.stats + ((int64_t)n * x.C + c) * 2;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float4 s = *reinterpret_cast<const float4*>(sp + 4 * h);
    const float m0 = s.x * invP, m1 = s.z * invP;
    const float v0 = s.y * invP - m0 * m0, v1 = s.w * invP - m1 * m1;
    const float r0 = gbv2::rsqrt_((v0 > 0.f ? v0 : 0.f) + p.eps), r1 = gbv2::rsqrt_((v1 > 0.f ? v1 : 0.f) + p.eps);
    k.a[2 * h] = r0;
    k.a[2 * h + 1] = r1;
    k.b[2 * h] = -m0 * r0;
    k.b[2 * h + 1] = -m1 * r1;
  }
}

// Phase 1 of thread `tid` of CTA `rank` of the cluster that owns (image n, channel group cgi): stream the CTA's pixel
// range once, fold the reflection border of the gradient, add the folded gradient to the residual gradient (RES),
// mask with the activation's derivative, stash (masked gradient, x) and return the thread's partial sums.
// st_g0 / st_g1 / st_x: this CTA's stash, [g.steps][THREADS] 16-byte vectors each.
template <bool RES, int U>
V2_HD void load_pass(const gb_in_bwd_params& p, const Geom& g, float neg_slope, int tid, int rank, int cgi, int n,
                     float4* st_g0, float4* st_g1, uint4* st_x, float (&s1)[8], float (&s2)[8]) {
  const gb_view& x = p.x;
  const gb_view& dy = p.dy_b;
  const int quad = tid % TPP, slot = tid / TPP;
  const int c = cgi * CG + quad * 8;
  const int W = x.W;
  const uint32_t P = (uint32_t)x.D * (uint32_t)x.H * (uint32_t)x.W;
  const uint32_t p0 = (uint32_t)rank * (uint32_t)g.ppc;
  const uint32_t p1 = (p0 + (uint32_t)g.ppc < P) ? p0 + (uint32_t)g.ppc : P;
  Consts k;
  load_consts(p, n, c, k);
#pragma unroll
  for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
  const float* gb = reinterpret_cast<const float*>(dy.ptr) + (int64_t)n * dy.sn + c;
  const uint16_t* xb = reinterpret_cast<const uint16_t*>(x.ptr) + (int64_t)n * x.sn + c;
  float* sb = RES ? reinterpret_cast<float*>(p.dy_sum.ptr) + (int64_t)n * p.dy_sum.sn + c : nullptr;
  const int gpad = dy.pad;
  const int gsx = (int)dy.sx, gsy = (int)dy.sy;
  for (int k0 = 0; k0 < g.steps; k0 += U) {
    float4 g0[U], g1[U], q0[U], q1[U];
    uint4 xv[U];
    int yy[U], xx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t pix = p0 + (uint32_t)((k0 + u) * SLOTS + slot);
      yy[u] = -1;
      if (k0 + u < g.steps && pix < p1) {
        const uint32_t y = gb_div(pix, g.divw);
        yy[u] = (int)y;
        xx[u] = (int)(pix - y * (uint32_t)W);
        const float* gp = gb + yy[u] * gsy + xx[u] * gsx;
        g0[u] = gbv2::ld_grad4(gp);
        g1[u] = gbv2::ld_grad4(gp + 4);
        xv[u] = gbv2::ld_bf16x8(xb + yy[u] * (int)x.sy + xx[u] * (int)x.sx);
        if (RES) {
          const float* rp = sb + yy[u] * (int)p.dy_sum.sy + xx[u] * (int)p.dy_sum.sx;
          q0[u] = gbv2::ld_own4(rp);
          q1[u] = gbv2::ld_own4(rp + 4);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (yy[u] < 0) continue;
      const int y = yy[u], px = xx[u];
      float gg[8] = {g0[u].x, g0[u].y, g0[u].z, g0[u].w, g1[u].x, g1[u].y, g1[u].z, g1[u].w};
      float xf[8];
      gbv2::unpack8(xv[u], xf);
      if (gpad > 0) {
        const int my = gbv2::mirror_of(y, dy.H, gpad), mx = gbv2::mirror_of(px, W, gpad);
        if (my != gbv2::NO_MIRROR) {
          const float* q = gb + my * gsy + px * gsx;
          gbv2::add4(gg, 0, gbv2::ld_grad4(q));
          gbv2::add4(gg, 4, gbv2::ld_grad4(q + 4));
        }
        if (mx != gbv2::NO_MIRROR) {
          const float* q = gb + y * gsy + mx * gsx;
          gbv2::add4(gg, 0, gbv2::ld_grad4(q));
          gbv2::add4(gg, 4, gbv2::ld_grad4(q + 4));
          if (my != gbv2::NO_MIRROR) {
            const float* qc = gb + my * gsy + mx * gsx;
            gbv2::add4(gg, 0, gbv2::ld_grad4(qc));
            gbv2::add4(gg, 4, gbv2::ld_grad4(qc + 4));
          }
        }
      }
      if (RES) {
        float* rp = sb + y * (int)p.dy_sum.sy + px * (int)p.dy_sum.sx;
        float4 o0, o1;
        o0.x = q0[u].x + gg[0]; o0.y = q0[u].y + gg[1]; o0.z = q0[u].z + gg[2]; o0.w = q0[u].w + gg[3];
        o1.x = q1[u].x + gg[4]; o1.y = q1[u].y + gg[5]; o1.z = q1[u].z + gg[6]; o1.w = q1[u].w + gg[7];
        *reinterpret_cast<float4*>(rp) = o0;
        *reinterpret_cast<float4*>(rp + 4) = o1;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = gbv2::fma_(xf[e], k.a[e], k.b[e]);
        if (!(xh > 0.f)) gg[e] *= neg_slope;
        s1[e] += gg[e];
        s2[e] = gbv2::fma_(gg[e], xh, s2[e]);
      }
      const int si = (k0 + u) * THREADS + tid;
      float4 w0, w1;
      w0.x = gg[0]; w0.y = gg[1]; w0.z = gg[2]; w0.w = gg[3];
      w1.x = gg[4]; w1.y = gg[5]; w1.z = gg[6]; w1.w = gg[7];
      st_g0[si] = w0;
      st_g1[si] = w1;
      st_x[si] = xv[u];
    }
  }
}

// Phase 2: dx from the stash and the cluster totals (tot1 = sum g, tot2 = sum g * xhat of this thread's 8 channels);
// db = the thread's partial sum of the fp32 dx (bias gradient).
V2_HD void apply_pass(const gb_in_bwd_params& p, const Geom& g, int tid, int rank, int cgi, int n, const float (&tot1)[8],
                      const float (&tot2)[8], const float4* st_g0, const float4* st_g1, const uint4* st_x, float (&db)[8]) {
  const gb_view& x = p.x;
  const int quad = tid % TPP, slot = tid / TPP;
  const int c = cgi * CG + quad * 8;
  const int W = x.W;
  const uint32_t P = (uint32_t)x.D * (uint32_t)x.H * (uint32_t)x.W;
  const uint32_t p0 = (uint32_t)rank * (uint32_t)g.ppc;
  const uint32_t p1 = (p0 + (uint32_t)g.ppc < P) ? p0 + (uint32_t)g.ppc : P;
  const float invP = 1.f / (float)P;
  Consts k;
  load_consts(p, n, c, k);
  float k1[8], k2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    k1[e] = k.a[e] * (tot1[e] * invP);
    k2[e] = k.a[e] * (tot2[e] * invP);
    db[e] = 0.f;
  }
  uint16_t* dxb = reinterpret_cast<uint16_t*>(p.dx.ptr) + (int64_t)n * p.dx.sn + c;
  for (int s = 0; s < g.steps; ++s) {
    const uint32_t pix = p0 + (uint32_t)(s * SLOTS + slot);
    if (pix >= p1) break;
    const uint32_t y = gb_div(pix, g.divw);
    const int px = (int)(pix - y * (uint32_t)W);
    const int si = s * THREADS + tid;
    const float4 w0 = st_g0[si], w1 = st_g1[si];
    const float ge[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    float xf[8], d[8];
    gbv2::unpack8(st_x[si], xf);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xh = gbv2::fma_(xf[e], k.a[e], k.b[e]);
      d[e] = gbv2::fma_(-xh, k2[e], gbv2::fma_(k.a[e], ge[e], -k1[e]));
      db[e] += d[e];
    }
    uint4 o;
    o.x = gbv2::pack2(d[0], d[1]);
    o.y = gbv2::pack2(d[2], d[3]);
    o.z = gbv2::pack2(d[4], d[5]);
    o.w = gbv2::pack2(d[6], d[7]);
    *reinterpret_cast<uint4*>(dxb + (int)y * (int)p.dx.sy + px * (int)p.dx.sx) = o;
  }
}

}  // namespace gbv3
