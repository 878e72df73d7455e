// Streaming fast paths of the fused InstanceNorm kernels (instnorm.cu keeps the fully general versions).
//
// Covered here: the patterns of the 2-D / 3-D generators and discriminators -- normalise + ReLU / LeakyReLU / none
// (+ residual added after the activation, + reflection border of the result) forward, and the matching backward
// with the gradient arriving on the (possibly reflection-padded) output domain, an optional residual-gradient
// accumulation and the bias-gradient column sum.  PReLU, residual-before-activation, scaled outputs and fp32
// destinations stay on the general kernels.
//
// Design: HBM-bound streaming.  Forward: a thread owns 8 consecutive channels (16 B of bf16), backward: 4 (8 B of
// bf16, 16 B of fp32) so that the per-channel constants of both passes fit in registers.  U pixels per thread are
// loaded back to back before any arithmetic (all loads independent), three 256-thread blocks per SM are resident
// and the grid is one resident wave.  A thread walks its pixels with a running (row, column) cursor -- the first
// version divided every pixel index by the row length and rebuilt 64-bit offsets per view, 28 instructions per
// element, and was issue-bound at 24 % occupancy (profiles/r01j) -- and the reflection border / mirrored gradient
// positions are a rarely taken branch.  Views are addressed as rows = D*H lines of W pixels (3-D views must be
// row-linear: sz == H*sy; bordered views are 2-D).  The backward reduction and apply passes run in ONE launch when
// the grid is co-resident: blocks publish their partial sums with atomics, meet at a grid barrier (cooperative
// launch guarantees co-residency) and re-read the tensors, which at the sizes where launch latency matters are
// still in L2.
#include "gb_common.cuh"
#include "gb_geometry.h"

namespace {

constexpr int FU = 4;        // forward: pixels in flight per thread (16 B + 16 B per pixel)
constexpr int BU = 4;        // backward: 16 B + 8 B (+ 16 B) per pixel
constexpr int FTHREADS = 256;
constexpr int FBLOCKS_PER_SM = 3;

struct FastGeom {
  int ppb;                   // pixels per block
  int nblocks;               // blocks per image
  int total_blocks;          // grid size (fused mode: barrier target)
};

struct Cursor {
  int y, x;                  // row (over D*H) and column of the thread's current pixel
};
__device__ __forceinline__ void advance(Cursor& c, int step, int W) {
  c.x += step;
  while (c.x >= W) {
    c.x -= W;
    ++c.y;
  }
}
// element offset of the cursor's pixel inside image n of view v (32-bit: tensors are checked < 2^31 elements)
__device__ __forceinline__ int off_of(const gb_view& v, const Cursor& c) { return c.y * (int)v.sy + c.x * (int)v.sx; }

__device__ __forceinline__ void unpack4(const uint2& u, float (&f)[4]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint2 pack4(const float (&f)[4]) {
  uint2 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  return o;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

// border index that reflects onto interior index i of an axis of length n with border p, or NO_MIRROR.  The host
// only takes this path when n > 2p + 1, so an index has at most one mirror image per axis.
constexpr int NO_MIRROR = -(1 << 20);
__device__ __forceinline__ int mirror_of(int i, int n, int p) {
  if (i >= 1 && i <= p) return -i;
  if (i <= n - 2 && i >= n - 1 - p) return 2 * (n - 1) - i;
  return NO_MIRROR;
}

// gradient on a reflection-padded domain folded onto interior pixel (y, x): adds the border positions that mirror
// onto it (the centre value is loaded by the caller); `img` points at channel c of image n
__device__ __forceinline__ void add_mirrors4(const gb_view& v, const float* img, int y, int x, int my, int mx, float (&f)[4]) {
  if (my != NO_MIRROR) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(img + my * (int)v.sy + x * (int)v.sx));
    f[0] += t.x; f[1] += t.y; f[2] += t.z; f[3] += t.w;
  }
  if (mx != NO_MIRROR) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(img + y * (int)v.sy + mx * (int)v.sx));
    f[0] += t.x; f[1] += t.y; f[2] += t.z; f[3] += t.w;
  }
  if (my != NO_MIRROR && mx != NO_MIRROR) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(img + my * (int)v.sy + mx * (int)v.sx));
    f[0] += t.x; f[1] += t.y; f[2] += t.z; f[3] += t.w;
  }
}

// store to every border position that reflects onto (y, x) (the centre is stored by the caller)
__device__ __forceinline__ void st8_mirrors(const gb_view& v, __nv_bfloat16* img, int y, int x, int my, int mx, const uint4& o) {
  if (my != NO_MIRROR) *reinterpret_cast<uint4*>(img + my * (int)v.sy + x * (int)v.sx) = o;
  if (mx != NO_MIRROR) *reinterpret_cast<uint4*>(img + y * (int)v.sy + mx * (int)v.sx) = o;
  if (my != NO_MIRROR && mx != NO_MIRROR) *reinterpret_cast<uint4*>(img + my * (int)v.sy + mx * (int)v.sx) = o;
}

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ forward
// y = act((x - mean) * rstd) [+ res], act(v) = v > 0 ? v : v * neg_slope  (neg_slope 1 = identity, 0 = ReLU)
template <bool RES>
__global__ void __launch_bounds__(FTHREADS, FBLOCKS_PER_SM)
in_fwd_fast_kernel(const __grid_constant__ gb_in_fwd_params p, const __grid_constant__ FastGeom g, float neg_slope) {
  gb_pdl_enter();
  const gb_view& x = p.x;
  const int C8 = x.C >> 3;
  const int slots = FTHREADS / C8;
  const int cg = threadIdx.x % C8;
  const int slot = threadIdx.x / C8;
  if (slot >= slots) return;
  const int c = cg * 8;
  const int n = blockIdx.y;
  const int W = x.W;
  const uint32_t P = (uint32_t)x.D * x.H * x.W;
  const uint32_t p0 = blockIdx.x * (uint32_t)g.ppb;
  const uint32_t p1 = min(P, p0 + (uint32_t)g.ppb);
  float mean[8], rstd[8];
  const float invP = 1.f / (float)P;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mean[e] = 0.f;
    rstd[e] = 1.f;
  }
  if (p.stats != nullptr) {
    const float4* sp = reinterpret_cast<const float4*>(p.stats + ((int64_t)n * x.C + c) * 2);
#pragma unroll
    for (int h = 0; h < 4; ++h) {  // (sum, sumsq) pairs of channels 2h, 2h+1
      const float4 a = sp[h];
      const float m0 = a.x * invP, m1 = a.z * invP;
      mean[2 * h] = m0;
      mean[2 * h + 1] = m1;
      rstd[2 * h] = rsqrtf(fmaxf(a.y * invP - m0 * m0, 0.f) + p.eps);
      rstd[2 * h + 1] = rsqrtf(fmaxf(a.w * invP - m1 * m1, 0.f) + p.eps);
    }
  }
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x.ptr) + (int64_t)n * x.sn + c;
  const __nv_bfloat16* rb = RES ? reinterpret_cast<const __nv_bfloat16*>(p.res.ptr) + (int64_t)n * p.res.sn + c : nullptr;
  __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(p.y.ptr) + (int64_t)n * p.y.sn + c;
  const int ypad = p.y.pad;
  Cursor cur;
  {
    const uint32_t pix = p0 + (uint32_t)slot;
    cur.y = (int)(pix / (uint32_t)W);
    cur.x = (int)(pix - (uint32_t)cur.y * (uint32_t)W);
  }
  for (uint32_t base = p0 + slot; base < p1; base += (uint32_t)(FU * slots)) {
    uint4 fx[FU], fr[FU];
    int cc[FU];  // (row << 16) | column of pixel u (extents are checked < 65536 on the host)
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      cc[u] = (cur.y << 16) | cur.x;
      if (base + (uint32_t)(u * slots) < p1) {
        fx[u] = __ldg(reinterpret_cast<const uint4*>(xb + off_of(x, cur)));
        if (RES) fr[u] = __ldg(reinterpret_cast<const uint4*>(rb + off_of(p.res, cur)));
      }
      advance(cur, slots, W);
    }
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      if (base + (uint32_t)(u * slots) < p1) {
        float f[8], r[8];
        unpack8(fx[u], f);
        if (RES) unpack8(fr[u], r);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v = (f[e] - mean[e]) * rstd[e];
          v = v > 0.f ? v : v * neg_slope;
          if (RES) v += r[e];
          f[e] = v;
        }
        const uint4 o = pack8(f);
        const Cursor at = {cc[u] >> 16, cc[u] & 0xFFFF};
        *reinterpret_cast<uint4*>(yb + off_of(p.y, at)) = o;
        if (ypad > 0) {
          const int my = mirror_of(at.y, p.y.H, ypad), mx = mirror_of(at.x, W, ypad);
          if (my != NO_MIRROR || mx != NO_MIRROR) st8_mirrors(p.y, yb, at.y, at.x, my, mx, o);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// g = dy (folded from the padded domain when dy.pad > 0);  residual gradient: dy_sum += g (unmasked);
// g *= (xhat > 0 ? 1 : neg_slope);  pass 0: bstats += (sum g, sum g*xhat);
// pass 1: dx = rstd * (g - mean(g) - xhat * mean(g*xhat)) -> bf16, dbias += column sums of the fp32 dx.
// PASS 0 / 1 = the two passes as separate launches, PASS 2 = both in one launch around a grid barrier.
template <bool RES>
__device__ __forceinline__ void in_bwd_fast_pass(const gb_in_bwd_params& p, const FastGeom& g, float neg_slope, int pass,
                                                 bool first_pass_of_fused, float* red) {
  constexpr int NU = RES ? 3 : BU;  // pixels in flight (the residual variant holds one more fp32 vector per pixel)
  const gb_view& x = p.x;
  const gb_view& dy = p.dy_b;
  const int C4 = x.C >> 2;
  const int slots = FTHREADS / C4;
  const int cg = threadIdx.x % C4;
  const int slot = threadIdx.x / C4;
  const int c = cg * 4;
  const int n = blockIdx.y;
  const int W = x.W;
  const uint32_t P = (uint32_t)x.D * x.H * x.W;
  const uint32_t p0 = blockIdx.x * (uint32_t)g.ppb;
  const uint32_t p1 = min(P, p0 + (uint32_t)g.ppb);
  const float invP = 1.f / (float)P;
  float mean[4], rstd[4], m1[4], m2[4], s1[4], s2[4];
  {
    const float4 a = *reinterpret_cast<const float4*>(p.stats + ((int64_t)n * x.C + c) * 2);
    const float4 b = *reinterpret_cast<const float4*>(p.stats + ((int64_t)n * x.C + c) * 2 + 4);
    const float s[4] = {a.x, a.z, b.x, b.z}, ss[4] = {a.y, a.w, b.y, b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float m = s[e] * invP;
      mean[e] = m;
      rstd[e] = rsqrtf(fmaxf(ss[e] * invP - m * m, 0.f) + p.eps);
      m1[e] = m2[e] = s1[e] = s2[e] = 0.f;
    }
  }
  if (pass == 1) {
    const float4 a = __ldcg(reinterpret_cast<const float4*>(p.bstats + ((int64_t)n * x.C + c) * 2));
    const float4 b = __ldcg(reinterpret_cast<const float4*>(p.bstats + ((int64_t)n * x.C + c) * 2 + 4));
    m1[0] = a.x * invP; m2[0] = a.y * invP; m1[1] = a.z * invP; m2[1] = a.w * invP;
    m1[2] = b.x * invP; m2[2] = b.y * invP; m1[3] = b.z * invP; m2[3] = b.w * invP;
  }
  const bool want_dbias = pass == 1 && p.dbias != nullptr;
  // the residual gradient is accumulated exactly once: in pass 0
  const bool do_res = RES && pass == 0;
  const float* gb = reinterpret_cast<const float*>(dy.ptr) + (int64_t)n * dy.sn + c;
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x.ptr) + (int64_t)n * x.sn + c;
  float* sb = RES ? reinterpret_cast<float*>(p.dy_sum.ptr) + (int64_t)n * p.dy_sum.sn + c : nullptr;
  __nv_bfloat16* db = reinterpret_cast<__nv_bfloat16*>(p.dx.ptr) + (int64_t)n * p.dx.sn + c;
  const int gpad = dy.pad;
  if (slot < slots) {
    Cursor cur;
    {
      const uint32_t pix = p0 + (uint32_t)slot;
      cur.y = (int)(pix / (uint32_t)W);
      cur.x = (int)(pix - (uint32_t)cur.y * (uint32_t)W);
    }
    for (uint32_t base = p0 + slot; base < p1; base += (uint32_t)(NU * slots)) {
      float4 lg[NU], lr[NU];
      uint2 lx[NU];
      int cc[NU];  // (row << 16) | column
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        cc[u] = (cur.y << 16) | cur.x;
        if (base + (uint32_t)(u * slots) < p1) {
          lg[u] = __ldcg(reinterpret_cast<const float4*>(gb + off_of(dy, cur)));
          lx[u] = __ldg(reinterpret_cast<const uint2*>(xb + off_of(x, cur)));
          if (do_res) lr[u] = __ldcg(reinterpret_cast<const float4*>(sb + off_of(p.dy_sum, cur)));
        }
        advance(cur, slots, W);
      }
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        if (base + (uint32_t)(u * slots) < p1) {
          float gg[4] = {lg[u].x, lg[u].y, lg[u].z, lg[u].w}, xv[4];
          unpack4(lx[u], xv);
          const Cursor at = {cc[u] >> 16, cc[u] & 0xFFFF};
          if (gpad > 0) {
            const int my = mirror_of(at.y, dy.H, gpad), mx = mirror_of(at.x, W, gpad);
            if (my != NO_MIRROR || mx != NO_MIRROR) add_mirrors4(dy, gb, at.y, at.x, my, mx, gg);
          }
          if (do_res)
            *reinterpret_cast<float4*>(sb + off_of(p.dy_sum, at)) =
                make_float4(lr[u].x + gg[0], lr[u].y + gg[1], lr[u].z + gg[2], lr[u].w + gg[3]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float xh = (xv[e] - mean[e]) * rstd[e];
            if (!(xh > 0.f)) gg[e] *= neg_slope;
            xv[e] = xh;
          }
          if (pass == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              s1[e] += gg[e];
              s2[e] += gg[e] * xv[e];
            }
          } else {
            float d[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) d[e] = rstd[e] * (gg[e] - m1[e] - xv[e] * m2[e]);
            *reinterpret_cast<uint2*>(db + off_of(p.dx, at)) = pack4(d);
            if (want_dbias) {  // bias gradient: sum of the fp32 dx (not of its bf16 rounding, see instnorm.cu)
#pragma unroll
              for (int e = 0; e < 4; ++e) s1[e] += d[e];
            }
          }
        }
      }
    }
  }
  // block reduction over the pixel slots, then one atomic per channel and block
  if (pass == 0 || want_dbias) {
    if (!first_pass_of_fused) __syncthreads();  // `red` may still be read by the previous pass
    if (slot < slots) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        red[(slot * x.C + c + e) * 2 + 0] = s1[e];
        red[(slot * x.C + c + e) * 2 + 1] = s2[e];
      }
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < x.C; ch += FTHREADS) {
      float a = 0.f, b = 0.f;
      for (int k = 0; k < slots; ++k) {
        a += red[(k * x.C + ch) * 2 + 0];
        b += red[(k * x.C + ch) * 2 + 1];
      }
      if (pass == 0) {
        atomicAdd(p.bstats + ((int64_t)n * x.C + ch) * 2 + 0, a);
        atomicAdd(p.bstats + ((int64_t)n * x.C + ch) * 2 + 1, b);
      } else {
        atomicAdd(p.dbias + ch, a);
      }
    }
  }
}

template <bool RES, int PASS>
__global__ void __launch_bounds__(FTHREADS, FBLOCKS_PER_SM)
in_bwd_fast_kernel(const __grid_constant__ gb_in_bwd_params p, const __grid_constant__ FastGeom g, float neg_slope) {
  gb_pdl_enter();
  extern __shared__ float red[];  // [slots][C][2]
  if (PASS == 2) {
    in_bwd_fast_pass<RES>(p, g, neg_slope, 0, true, red);
    unsigned int* counter = reinterpret_cast<unsigned int*>(p.bstats + (int64_t)p.x.N * p.x.C * 2);
    grid_barrier(counter, (unsigned int)g.total_blocks);
    in_bwd_fast_pass<RES>(p, g, neg_slope, 1, false, red);
  } else {
    in_bwd_fast_pass<RES>(p, g, neg_slope, PASS, true, red);
  }
}

// ------------------------------------------------------------------------------------------------ backward, lean
// Third generation of the fused backward (knob 22 = 4, then the default): the same two passes around a grid barrier,
// written for instruction count -- the first generation is ISSUE-bound, not memory-bound (29 instructions per element
// and pass, identical time with a cold and a warm L2: profiles/r02b_in_microbench_b8.txt).  Eight channels per thread
// (16 B of bf16 x, 2 x 16 B of fp32 dy: per-pixel index arithmetic is shared by twice as many elements), U pixels in
// flight per thread with all their loads issued first, the mirrored
// border positions of a reflection-padded gradient behind one unsigned compare per axis.
constexpr int LTHREADS = 256;
constexpr int LBLOCKS_PER_SM = 2;

__device__ __forceinline__ bool near_edge(int i, int n, int g) {
  // true for the interior indices that have a mirror image in a border of width g: 1..g and n-1-g..n-2
  return (unsigned)(i - 1) < (unsigned)g || (unsigned)(i - (n - 1 - g)) < (unsigned)g;
}

__device__ __forceinline__ void add_mirrors8(const gb_view& v, const float* img, int y, int x, int H, int W, int g,
                                             float (&f)[8]) {
  const int my = mirror_of(y, H, g), mx = mirror_of(x, W, g);
  auto add = [&](const float* q) {
    const float4 a = __ldcg(reinterpret_cast<const float4*>(q)), b = __ldcg(reinterpret_cast<const float4*>(q) + 1);
    f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
  };
  if (my != NO_MIRROR) add(img + my * (int)v.sy + x * (int)v.sx);
  if (mx != NO_MIRROR) add(img + y * (int)v.sy + mx * (int)v.sx);
  if (my != NO_MIRROR && mx != NO_MIRROR) add(img + my * (int)v.sy + mx * (int)v.sx);
}

template <bool RES, int PASS>
__device__ __forceinline__ void in_bwd_lean_pass(const gb_in_bwd_params& p, const FastGeom& g, float neg_slope, float* red,
                                                 bool sync_first) {
  constexpr int U = RES && PASS == 0 ? 2 : 4;
  const gb_view& x = p.x;
  const gb_view& dy = p.dy_b;
  const int C8 = x.C >> 3;
  const int slots = LTHREADS / C8;
  const int cg = threadIdx.x % C8;
  const int slot = threadIdx.x / C8;
  const int c = cg * 8;
  const int n = blockIdx.y;
  const int W = x.W, H = dy.H;
  const uint32_t P = (uint32_t)x.D * x.H * x.W;
  const uint32_t p0 = blockIdx.x * (uint32_t)g.ppb;
  const uint32_t p1 = min(P, p0 + (uint32_t)g.ppb);
  const float invP = 1.f / (float)P;
  float ka[8], kb[8], m1[8], m2[8], s1[8], s2[8];   // x-hat = (x - kb) * ka: the FORWARD kernel's expression, so that
                                                    // the activation mask is the one the forward pass applied
  {
    const float4* sp = reinterpret_cast<const float4*>(p.stats + ((int64_t)n * x.C + c) * 2);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float4 a = sp[h];
      const float mu0 = a.x * invP, mu1 = a.z * invP;
      const float r0 = rsqrtf(fmaxf(a.y * invP - mu0 * mu0, 0.f) + p.eps);
      const float r1 = rsqrtf(fmaxf(a.w * invP - mu1 * mu1, 0.f) + p.eps);
      ka[2 * h] = r0; kb[2 * h] = mu0;
      ka[2 * h + 1] = r1; kb[2 * h + 1] = mu1;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) m1[e] = m2[e] = s1[e] = s2[e] = 0.f;
  if (PASS == 1) {
    const float4* bp = reinterpret_cast<const float4*>(p.bstats + ((int64_t)n * x.C + c) * 2);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float4 a = __ldcg(bp + h);
      m1[2 * h] = a.x * invP; m2[2 * h] = a.y * invP; m1[2 * h + 1] = a.z * invP; m2[2 * h + 1] = a.w * invP;
    }
  }
  const bool want_dbias = PASS == 1 && p.dbias != nullptr;
  const float* gb = reinterpret_cast<const float*>(dy.ptr) + (int64_t)n * dy.sn + c;
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x.ptr) + (int64_t)n * x.sn + c;
  float* sb = RES ? reinterpret_cast<float*>(p.dy_sum.ptr) + (int64_t)n * p.dy_sum.sn + c : nullptr;
  __nv_bfloat16* db = reinterpret_cast<__nv_bfloat16*>(p.dx.ptr) + (int64_t)n * p.dx.sn + c;
  const int gpad = dy.pad;
  const int gsy = (int)dy.sy, gsx = (int)dy.sx, xsy = (int)x.sy, xsx = (int)x.sx;
  if (slot < slots) {
    int cy, cx;
    {
      const uint32_t pix = p0 + (uint32_t)slot;
      cy = (int)(pix / (uint32_t)W);
      cx = (int)(pix - (uint32_t)cy * (uint32_t)W);
    }
    for (uint32_t base = p0 + slot; base < p1; base += (uint32_t)(U * slots)) {
      float4 ga[U], gb2[U];
      uint4 lx[U];
      int py[U], px[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        py[u] = cy;
        px[u] = cx;
        if (base + (uint32_t)(u * slots) < p1) {
          const float4* gp = reinterpret_cast<const float4*>(gb + cy * gsy + cx * gsx);
          ga[u] = __ldcg(gp);
          gb2[u] = __ldcg(gp + 1);
          lx[u] = __ldg(reinterpret_cast<const uint4*>(xb + cy * xsy + cx * xsx));
        }
        cx += slots;
        while (cx >= W) {
          cx -= W;
          ++cy;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (base + (uint32_t)(u * slots) < p1) {
          float gg[8] = {ga[u].x, ga[u].y, ga[u].z, ga[u].w, gb2[u].x, gb2[u].y, gb2[u].z, gb2[u].w};
          float xv[8];
          unpack8(lx[u], xv);
          if (gpad > 0 && (near_edge(py[u], H, gpad) || near_edge(px[u], W, gpad)))
            add_mirrors8(dy, gb, py[u], px[u], H, W, gpad, gg);
          if (RES && PASS == 0) {
            float4* rp = reinterpret_cast<float4*>(sb + py[u] * (int)p.dy_sum.sy + px[u] * (int)p.dy_sum.sx);
            float4 a = __ldcg(rp), b = __ldcg(rp + 1);
            a.x += gg[0]; a.y += gg[1]; a.z += gg[2]; a.w += gg[3];
            b.x += gg[4]; b.y += gg[5]; b.z += gg[6]; b.w += gg[7];
            rp[0] = a;
            rp[1] = b;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float xh = (xv[e] - kb[e]) * ka[e];
            gg[e] = xh > 0.f ? gg[e] : gg[e] * neg_slope;
            xv[e] = xh;
          }
          if (PASS == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              s1[e] += gg[e];
              s2[e] = fmaf(gg[e], xv[e], s2[e]);
            }
          } else {
            float d[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) d[e] = ka[e] * (gg[e] - m1[e] - xv[e] * m2[e]);
            *reinterpret_cast<uint4*>(db + py[u] * (int)p.dx.sy + px[u] * (int)p.dx.sx) = pack8(d);
            if (want_dbias) {
#pragma unroll
              for (int e = 0; e < 8; ++e) s1[e] += d[e];
            }
          }
        }
      }
    }
  }
  // block reduction over the pixel slots, then one atomic per channel and block
  if (PASS == 0 || want_dbias) {
    if (sync_first) __syncthreads();  // `red` may still be read by the previous pass
    if (slot < slots) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        red[(slot * x.C + c + e) * 2 + 0] = s1[e];
        red[(slot * x.C + c + e) * 2 + 1] = s2[e];
      }
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < x.C; ch += LTHREADS) {
      float a = 0.f, b = 0.f;
      for (int k = 0; k < slots; ++k) {
        a += red[(k * x.C + ch) * 2 + 0];
        b += red[(k * x.C + ch) * 2 + 1];
      }
      if (PASS == 0) {
        atomicAdd(p.bstats + ((int64_t)n * x.C + ch) * 2 + 0, a);
        atomicAdd(p.bstats + ((int64_t)n * x.C + ch) * 2 + 1, b);
      } else {
        atomicAdd(p.dbias + ch, a);
      }
    }
  }
}

template <bool RES, int PASS>
__global__ void __launch_bounds__(LTHREADS, LBLOCKS_PER_SM)
in_bwd_lean_kernel(const __grid_constant__ gb_in_bwd_params p, const __grid_constant__ FastGeom g, float neg_slope) {
  gb_pdl_enter();
  extern __shared__ float red[];  // [slots][C][2]
  if (PASS == 2) {
    in_bwd_lean_pass<RES, 0>(p, g, neg_slope, red, false);
    unsigned int* counter = reinterpret_cast<unsigned int*>(p.bstats + (int64_t)p.x.N * p.x.C * 2);
    grid_barrier(counter, (unsigned int)g.total_blocks);
    in_bwd_lean_pass<RES, 1>(p, g, neg_slope, red, true);
  } else if (PASS == 0) {
    in_bwd_lean_pass<RES, 0>(p, g, neg_slope, red, false);
  } else {
    in_bwd_lean_pass<RES, 1>(p, g, neg_slope, red, false);
  }
}

// ------------------------------------------------------------------------------------------------ host side
// rows = D*H lines of W pixels with one row stride: any 2-D view, 3-D views whose planes are row-contiguous
bool row_addressable(const gb_view& v) { return v.D == 1 || (v.pad == 0 && v.sz == (int64_t)v.H * v.sy); }
bool small_offsets(const gb_view& v) {  // 32-bit in-image offsets, 16-bit packed (row, column)
  return ((int64_t)v.D * v.H + 2 * v.pad) * v.sy + (int64_t)(v.W + 2 * v.pad) * v.sx < (1ll << 31) &&
         (int64_t)v.D * v.H < 32768 && v.W < 65536;
}
bool aligned(const gb_view& v, int elem_bytes, int vec) {
  // vec-channel vectors: every pixel's channel-slice start is aligned to vec elements
  return ((uintptr_t)v.ptr % (elem_bytes * vec)) == 0 && v.sx % vec == 0 && v.sy % vec == 0 && v.sz % vec == 0 &&
         v.sn % vec == 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// one resident wave: `cap` co-resident blocks shared by the N images
FastGeom plan(const gb_view& x, int cap, bool* fits, int vec) {
  FastGeom g;
  const int C4 = x.C / vec;
  const int slots = FTHREADS / C4;
  const int64_t P = (int64_t)x.D * x.H * x.W;
  int nb = cap / x.N;
  *fits = nb >= 1;
  if (nb < 1) nb = 1;
  int64_t ppb = (P + nb - 1) / nb;
  ppb = (ppb + slots - 1) / slots * slots;
  if (ppb < slots) ppb = slots;
  g.ppb = (int)ppb;
  g.nblocks = (int)((P + ppb - 1) / ppb);
  g.total_blocks = g.nblocks * x.N;
  return g;
}

template <bool RES>
int launch_fwd(const gb_in_fwd_params& p, float neg_slope, cudaStream_t st) {
  bool fits;
  const FastGeom g = plan(p.x, num_sms() * FBLOCKS_PER_SM, &fits, 8);
  gb_klaunch(in_fwd_fast_kernel<RES>, dim3(g.nblocks, p.x.N), FTHREADS, 0, st, p, g, neg_slope);
  GB_LAUNCH_CHECK();
  return 0;
}

template <bool RES>
int launch_bwd(const gb_in_bwd_params& p, float neg_slope, cudaStream_t st) {
  const int C4 = p.x.C / 4;
  const int slots = FTHREADS / C4;
  const size_t smem = sizeof(float) * 2 * slots * p.x.C;
  static int occ = -1;  // co-resident blocks per SM of the single-launch kernel (per instantiation)
  static size_t occ_smem = 0;
  if (occ < 0 || occ_smem != smem) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, in_bwd_fast_kernel<RES, 2>, FTHREADS, smem) != cudaSuccess) o = 0;
    cudaGetLastError();
    occ = o;
    occ_smem = smem;
  }
  bool fits = false;
  FastGeom g = plan(p.x, num_sms() * (occ > 0 ? occ : FBLOCKS_PER_SM), &fits, 4);
  const dim3 grid(g.nblocks, p.x.N);
  // one cooperative launch around a grid barrier only on request (knob 6 = 2): it needs the whole grid co-resident, so
  // it cannot overlap a kernel of another stream; as two launches the step is 3.7 % faster under two-stream execution
  // (r02n: 372.5 -> 386.1 img/s) and the same single-stream
  if (occ > 0 && fits && g_gb_knobs[6] == 2) {
    void* args[] = {(void*)&p, (void*)&g, (void*)&neg_slope};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)in_bwd_fast_kernel<RES, 2>, grid, dim3(FTHREADS), args,
                                                smem, st);
    if (e == cudaSuccess) {
      __atomic_fetch_add(&g_gb_launches, 1ull, __ATOMIC_RELAXED);
      return 0;
    }
    cudaGetLastError();  // cooperative launch not possible here: fall through to two launches
  }
  gb_klaunch(in_bwd_fast_kernel<RES, 0>, grid, FTHREADS, smem, st, p, g, neg_slope);
  GB_LAUNCH_CHECK();
  gb_klaunch(in_bwd_fast_kernel<RES, 1>, grid, FTHREADS, smem, st, p, g, neg_slope);
  GB_LAUNCH_CHECK();
  return 0;
}

template <bool RES>
int launch_bwd_lean(const gb_in_bwd_params& p, float neg_slope, cudaStream_t st) {
  const int C8 = p.x.C / 8;
  const int slots = LTHREADS / C8;
  const size_t smem = sizeof(float) * 2 * slots * p.x.C;
  static int occ = -1;  // co-resident blocks per SM of the single-launch kernel (per instantiation)
  static size_t occ_smem = 0;
  if (occ < 0 || occ_smem != smem) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, in_bwd_lean_kernel<RES, 2>, LTHREADS, smem) != cudaSuccess) o = 0;
    cudaGetLastError();
    occ = o;
    occ_smem = smem;
  }
  bool fits = false;
  FastGeom g = plan(p.x, num_sms() * (occ > 0 ? occ : LBLOCKS_PER_SM), &fits, 8);
  const dim3 grid(g.nblocks, p.x.N);
  // one cooperative launch around a grid barrier only on request (knob 6 = 2): it needs the whole grid co-resident, so
  // it cannot overlap a kernel of another stream; as two launches the step is 3.7 % faster under two-stream execution
  // (r02n: 372.5 -> 386.1 img/s) and the same single-stream
  if (occ > 0 && fits && g_gb_knobs[6] == 2) {
    void* args[] = {(void*)&p, (void*)&g, (void*)&neg_slope};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)in_bwd_lean_kernel<RES, 2>, grid, dim3(LTHREADS), args, smem, st);
    if (e == cudaSuccess) {
      __atomic_fetch_add(&g_gb_launches, 1ull, __ATOMIC_RELAXED);
      return 0;
    }
    cudaGetLastError();  // cooperative launch not possible here: fall through to two launches
  }
  gb_klaunch(in_bwd_lean_kernel<RES, 0>, grid, LTHREADS, smem, st, p, g, neg_slope);
  GB_LAUNCH_CHECK();
  gb_klaunch(in_bwd_lean_kernel<RES, 1>, grid, LTHREADS, smem, st, p, g, neg_slope);
  GB_LAUNCH_CHECK();
  return 0;
}

bool act_to_slope(int act, float slope, float* out) {
  switch (act) {
    case GB_ACT_NONE: *out = 1.f; return true;
    case GB_ACT_RELU: *out = 0.f; return true;
    case GB_ACT_LEAKY: *out = slope; return true;
  }
  return false;
}

}  // namespace

// -1: not covered (caller uses the general kernel), 0: launched, >0: error
int gb_in_fwd_fast(const gb_in_fwd_params& p, cudaStream_t st) {
  if (g_gb_knobs[7] != 0) return -1;
  float ns;
  if (!act_to_slope(p.act, p.act_slope, &ns)) return -1;
  if (p.res_before_act || (p.out_scale != 0.f && p.out_scale != 1.f)) return -1;
  const gb_view& x = p.x;
  if (x.C % 8 != 0 || x.C / 8 > FTHREADS || (int64_t)x.D * x.H * x.W >= (1ll << 31)) return -1;
  const bool has_res = p.res.ptr != nullptr;
  if (!aligned(x, 2, 8) || !aligned(p.y, 2, 8) || (has_res && !aligned(p.res, 2, 8))) return -1;
  if (p.stats != nullptr && (uintptr_t)p.stats % 16 != 0) return -1;
  if (!row_addressable(x) || !row_addressable(p.y) || (has_res && !row_addressable(p.res))) return -1;
  if (!small_offsets(x) || !small_offsets(p.y) || (has_res && !small_offsets(p.res))) return -1;
  // (x and res may be interior views of bordered buffers: only their interior is read)
  if (p.y.pad > 0 && (p.y.D != 1 || p.y.H <= 2 * p.y.pad + 1 || p.y.W <= 2 * p.y.pad + 1)) return -1;
  return has_res ? launch_fwd<true>(p, ns, st) : launch_fwd<false>(p, ns, st);
}

int gb_in_bwd_fast(const gb_in_bwd_params& p, cudaStream_t st) {
  if (g_gb_knobs[7] != 0) return -1;
  float ns;
  if (!act_to_slope(p.act, p.act_slope, &ns)) return -1;
  if (p.stats == nullptr || p.bstats == nullptr) return -1;
  if (p.dy_a.ptr != nullptr || p.dy_b.ptr == nullptr) return -1;
  if (p.res_before_act || p.dx_fp32_acc || p.dprelu != nullptr) return -1;
  if (p.out_scale != 0.f && p.out_scale != 1.f) return -1;
  const bool has_res = p.dy_sum.ptr != nullptr;
  if (has_res && !p.dy_sum_acc) return -1;
  const gb_view& x = p.x;
  if (x.C % 4 != 0 || x.C / 4 > FTHREADS || (int64_t)x.D * x.H * x.W >= (1ll << 31)) return -1;
  if (!aligned(x, 2, 4) || !aligned(p.dx, 2, 4) || !aligned(p.dy_b, 4, 4) || (has_res && !aligned(p.dy_sum, 4, 4))) return -1;
  if ((uintptr_t)p.stats % 16 != 0 || (uintptr_t)p.bstats % 16 != 0) return -1;
  if (!row_addressable(x) || !row_addressable(p.dx) || !row_addressable(p.dy_b) || (has_res && !row_addressable(p.dy_sum)))
    return -1;
  if (!small_offsets(x) || !small_offsets(p.dx) || !small_offsets(p.dy_b) || (has_res && !small_offsets(p.dy_sum))) return -1;
  // (x, dx and dy_sum may be interior views of bordered buffers: only their interior is touched)
  if (p.dy_b.pad > 0 && (p.dy_b.D != 1 || p.dy_b.H <= 2 * p.dy_b.pad + 1 || p.dy_b.W <= 2 * p.dy_b.pad + 1)) return -1;
  // third generation (8 channels per thread): knob 22 = 4 opts in, 5 opts out once it is the default
  const int C8 = x.C / 8;
  const bool lean_ok = x.C % 8 == 0 && C8 <= LTHREADS && LTHREADS % C8 == 0 && aligned(x, 2, 8) && aligned(p.dx, 2, 8) &&
                       aligned(p.dy_b, 4, 8) && (!has_res || aligned(p.dy_sum, 4, 8));
  if (lean_ok && g_gb_knobs[22] == 4) {
    ++g_gb_knobs[23];
    return has_res ? launch_bwd_lean<true>(p, ns, st) : launch_bwd_lean<false>(p, ns, st);
  }
  return has_res ? launch_bwd<true>(p, ns, st) : launch_bwd<false>(p, ns, st);
}
