// InstanceNorm (affine=False) fused with the following activation / residual add, forward and backward,
// on channels-last bf16 views.  HBM-bound: every kernel moves 16 bytes per thread per pixel, a warp covers
// consecutive channels of the same pixel (coalesced 128..512 B), statistics are fp32.
//
// Forward:  gb_in_stats (sum, sum^2 per (n,c))  ->  gb_in_fwd (normalise + act + residual, also writes the
//           reflection border of the destination so the next convolution needs no padding pass).
// Backward: reduce (sum g, sum g*xhat) -> apply dx = rstd * (g - mean(g) - xhat * mean(g*xhat)).
//           The incoming gradient may arrive on the padded domain (dgrad of a reflection-padded conv); the
//           border is folded back onto the interior while it is read.
#include "gb_common.cuh"
#include "gb_geometry.h"

namespace {

struct Pix {
  int z, y, x;
};
__device__ __forceinline__ Pix decode_pix(int64_t pix64, const gb_view& v) {
  Pix r;  // 32-bit on purpose (pixel counts are checked < 2^31 on the host)
  const uint32_t pix = (uint32_t)pix64, W = (uint32_t)v.W, H = (uint32_t)v.H;
  const uint32_t t = pix / W;
  r.x = (int)(pix - t * W);
  r.z = (int)(t / H);
  r.y = (int)(t - (uint32_t)r.z * H);
  return r;
}
__device__ __forceinline__ void load8(const gb_view& v, int n, int z, int y, int x, int cg, float (&f)[8]) {
  const __nv_bfloat16* ptr = reinterpret_cast<const __nv_bfloat16*>(v.ptr);
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(ptr + gb_pix_offset(v, n, z, y, x) + cg * 8));
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
// fp32 view variants (activation gradients)
__device__ __forceinline__ void load8f(const gb_view& v, int n, int z, int y, int x, int cg, float (&f)[8]) {
  const float* ptr = reinterpret_cast<const float*>(v.ptr) + gb_pix_offset(v, n, z, y, x) + cg * 8;
  const float4 a = __ldg(reinterpret_cast<const float4*>(ptr)), b = __ldg(reinterpret_cast<const float4*>(ptr) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8f(const gb_view& v, int n, int z, int y, int x, int cg, const float (&f)[8]) {
  float* ptr = reinterpret_cast<float*>(v.ptr) + gb_pix_offset(v, n, z, y, x) + cg * 8;
  reinterpret_cast<float4*>(ptr)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(ptr)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}
__device__ __forceinline__ void store8(const gb_view& v, int n, int z, int y, int x, int cg, const uint4& o) {
  __nv_bfloat16* ptr = reinterpret_cast<__nv_bfloat16*>(v.ptr);
  *reinterpret_cast<uint4*>(ptr + gb_pix_offset(v, n, z, y, x) + cg * 8) = o;
}
// store to (y,x) and to every border position that reflects onto it
__device__ __forceinline__ void store8_reflect(const gb_view& v, int n, int z, int y, int x, int cg, const uint4& o) {
  const int p = v.pad;
  int ys[3], xs[3], ny = 1, nx = 1;
  ys[0] = y;
  xs[0] = x;
  if (p > 0) {
    if (y >= 1 && y <= p) ys[ny++] = -y;
    if (y <= v.H - 2 && y >= v.H - 1 - p) ys[ny++] = 2 * (v.H - 1) - y;
    if (x >= 1 && x <= p) xs[nx++] = -x;
    if (x <= v.W - 2 && x >= v.W - 1 - p) xs[nx++] = 2 * (v.W - 1) - x;
  }
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) store8(v, n, z, ys[a], xs[b], cg, o);
}
// gradient on the padded domain folded back to interior pixel (y,x): the centre value is loaded by the caller
// (straight-line, batched across pixels); this adds the border positions that reflect onto (y,x) -- only pixels
// within `pad` of an edge have any
__device__ __forceinline__ void add_mirrors(const gb_view& v, int n, int z, int y, int x, int cg, float (&f)[8]) {
  const int p = v.pad;
  if (p == 0) return;
  const bool ynear = (y >= 1 && y <= p) || (y <= v.H - 2 && y >= v.H - 1 - p);
  const bool xnear = (x >= 1 && x <= p) || (x <= v.W - 2 && x >= v.W - 1 - p);
  if (!ynear && !xnear) return;
  int ys[3], xs[3], ny = 1, nx = 1;
  ys[0] = y;
  xs[0] = x;
  if (y >= 1 && y <= p) ys[ny++] = -y;
  if (y <= v.H - 2 && y >= v.H - 1 - p) ys[ny++] = 2 * (v.H - 1) - y;
  if (x >= 1 && x <= p) xs[nx++] = -x;
  if (x <= v.W - 2 && x >= v.W - 1 - p) xs[nx++] = 2 * (v.W - 1) - x;
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) {
      if (a == 0 && b == 0) continue;  // centre already loaded
      float t[8];
      load8f(v, n, z, ys[a], xs[b], cg, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += t[e];
    }
}

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  switch (act) {
    case GB_ACT_RELU: return fmaxf(v, 0.f);
    case GB_ACT_LEAKY:
    case GB_ACT_PRELU: return v > 0.f ? v : v * slope;
    case GB_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// ------------------------------------------------------------------------------------------- statistics
__global__ void in_stats_kernel(gb_view x, float* __restrict__ stats, int pix_per_block) {
  gb_pdl_enter();
  extern __shared__ float red[];  // [slots][2C]
  const int C8 = x.C >> 3;
  const int slots = blockDim.x / C8;
  const int cg = threadIdx.x % C8;
  const int slot = threadIdx.x / C8;
  const int n = blockIdx.y;
  const int64_t P = (int64_t)x.D * x.H * x.W;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p1 = min(P, p0 + pix_per_block);
  float s[8], ss[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = ss[e] = 0.f;
  if (slot < slots) {
    constexpr int U = 4;  // pixels in flight per thread (the kernel is latency bound otherwise)
    for (int64_t pix = p0 + slot; pix < p1; pix += (int64_t)U * slots) {
      float f[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t pu = pix + (int64_t)u * slots;
        if (pu < p1) {
          const Pix q = decode_pix(pu, x);
          load8(x, n, q.z, q.y, q.x, cg, f[u]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[u][e] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s[e] += f[u][e];
          ss[e] += f[u][e] * f[u][e];
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[(slot * x.C + cg * 8 + e) * 2 + 0] = s[e];
      red[(slot * x.C + cg * 8 + e) * 2 + 1] = ss[e];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * x.C; c += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < slots; ++k) t += red[k * 2 * x.C + c];
    atomicAdd(stats + (int64_t)n * 2 * x.C + c, t);
  }
}

// ------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(256, 2) in_fwd_kernel(const __grid_constant__ gb_in_fwd_params p, int pix_per_block) {
  gb_pdl_enter();
  const gb_view& x = p.x;
  const int C8 = x.C >> 3;
  const int slots = blockDim.x / C8;
  const int cg = threadIdx.x % C8;
  const int slot = threadIdx.x / C8;
  if (slot >= slots) return;
  const int n = blockIdx.y;
  const int64_t P = (int64_t)x.D * x.H * x.W;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p1 = min(P, p0 + pix_per_block);
  float mean[8], rstd[8], slope[8];
  const float invP = 1.f / (float)P;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = cg * 8 + e;
    if (p.stats) {
      const float s = p.stats[((int64_t)n * x.C + c) * 2], ss = p.stats[((int64_t)n * x.C + c) * 2 + 1];
      const float m = s * invP;
      const float var = fmaxf(ss * invP - m * m, 0.f);
      mean[e] = m;
      rstd[e] = rsqrtf(var + p.eps);
    } else {
      mean[e] = 0.f;
      rstd[e] = 1.f;
    }
    slope[e] = (p.act == GB_ACT_PRELU) ? p.prelu[c] : p.act_slope;
  }
  const bool has_res = p.res.ptr != nullptr;
  const float oscale = p.out_scale == 0.f ? 1.f : p.out_scale;
  constexpr int U = 2;  // pixels in flight per thread
  for (int64_t pix = p0 + slot; pix < p1; pix += (int64_t)U * slots) {
    float f[U][8], r[U][8];
    Pix q[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pu = pix + (int64_t)u * slots;
      ok[u] = pu < p1;
      if (ok[u]) {
        q[u] = decode_pix(pu, x);
        load8(x, n, q[u].z, q[u].y, q[u].x, cg, f[u]);
        if (has_res) load8(p.res, n, q[u].z, q[u].y, q[u].x, cg, r[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = (f[u][e] - mean[e]) * rstd[e];
        if (has_res && p.res_before_act) v += r[u][e];
        v = act_fwd(v, p.act, slope[e]) * oscale;
        if (has_res && !p.res_before_act) v += r[u][e];
        f[u][e] = v;
      }
      const uint4 o = pack8(f[u]);
      if (p.y.pad > 0) store8_reflect(p.y, n, q[u].z, q[u].y, q[u].x, cg, o);
      else store8(p.y, n, q[u].z, q[u].y, q[u].x, cg, o);
    }
  }
}

// ------------------------------------------------------------------------------------------- backward
// g = act'(.) * (dy_a + fold(dy_b));  MODE 0: reduce (sum g, sum g*xhat), optionally write dy_sum
//                                     MODE 1: apply dx = rstd * (g - m1 - xhat*m2)
template <int MODE>
__global__ void __launch_bounds__(256, 2) in_bwd_kernel(const __grid_constant__ gb_in_bwd_params p, int pix_per_block) {
  gb_pdl_enter();
  extern __shared__ float red[];
  const gb_view& x = p.x;
  const int C8 = x.C >> 3;
  const int slots = blockDim.x / C8;
  const int cg = threadIdx.x % C8;
  const int slot = threadIdx.x / C8;
  const int n = blockIdx.y;
  const int64_t P = (int64_t)x.D * x.H * x.W;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p1 = min(P, p0 + pix_per_block);
  const bool norm = p.stats != nullptr;
  float mean[8], rstd[8], slope[8], m1[8], m2[8], s1[8], s2[8], sp[8];
  const bool want_dbias = MODE == 1 && p.dbias != nullptr;
  const bool rba = p.res_before_act != 0 && p.res.ptr != nullptr;
  const float oscale = p.out_scale == 0.f ? 1.f : p.out_scale;
  // PReLU slope gradient without normalisation has no reduce pass: it is reduced in the apply pass
  const bool want_dprelu1 = MODE == 1 && p.stats == nullptr && p.act == GB_ACT_PRELU && p.dprelu != nullptr;
  const float invP = 1.f / (float)P;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = cg * 8 + e;
    mean[e] = 0.f; rstd[e] = 1.f; m1[e] = 0.f; m2[e] = 0.f; s1[e] = 0.f; s2[e] = 0.f; sp[e] = 0.f;
    if (norm) {
      const float s = p.stats[((int64_t)n * x.C + c) * 2], ss = p.stats[((int64_t)n * x.C + c) * 2 + 1];
      const float m = s * invP;
      mean[e] = m;
      rstd[e] = rsqrtf(fmaxf(ss * invP - m * m, 0.f) + p.eps);
      if (MODE == 1) {
        m1[e] = p.bstats[((int64_t)n * x.C + c) * 2] * invP;
        m2[e] = p.bstats[((int64_t)n * x.C + c) * 2 + 1] * invP;
      }
    }
    slope[e] = (p.act == GB_ACT_PRELU) ? p.prelu[c] : p.act_slope;
  }
  const bool has_a = p.dy_a.ptr != nullptr, has_b = p.dy_b.ptr != nullptr;
  // reduce pass already materialised a+fold(b) (only when dy_sum holds exactly that: no accumulate, not masked)
  const bool use_sum = MODE == 1 && p.dy_sum.ptr != nullptr && norm && !p.dy_sum_acc && !p.res_before_act;
  if (slot < slots) {
    constexpr int U = 2;  // pixels in flight per thread: all loads of both pixels are issued before any math
    for (int64_t pix0 = p0 + slot; pix0 < p1; pix0 += (int64_t)U * slots) {
      float gU[U][8], xU[U][8];
      Pix qU[U];
      bool okU[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t pu = pix0 + (int64_t)u * slots;
        okU[u] = pu < p1;
        if (!okU[u]) continue;
        const Pix q = decode_pix(pu, x);
        qU[u] = q;
        if (use_sum) {
          load8f(p.dy_sum, n, q.z, q.y, q.x, cg, gU[u]);
        } else {
          if (has_b) load8f(p.dy_b, n, q.z, q.y, q.x, cg, gU[u]);
          else load8f(p.dy_a, n, q.z, q.y, q.x, cg, gU[u]);
        }
        if (norm || p.y.ptr == nullptr) {
          if (norm || p.act != GB_ACT_NONE) load8(x, n, q.z, q.y, q.x, cg, xU[u]);
        } else if (p.act != GB_ACT_NONE) {
          load8(p.y, n, q.z, q.y, q.x, cg, xU[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!okU[u]) continue;
        const Pix q = qU[u];
        float (&g)[8] = gU[u];
        float (&xv)[8] = xU[u];
        if (!use_sum) {
          if (has_b) {
            add_mirrors(p.dy_b, n, q.z, q.y, q.x, cg, g);
            if (has_a) {
              float t[8];
              load8f(p.dy_a, n, q.z, q.y, q.x, cg, t);
#pragma unroll
              for (int e = 0; e < 8; ++e) g[e] += t[e];
            }
          }
          // residual added AFTER the activation: its gradient is the unmasked total
          if (!rba && p.dy_sum.ptr != nullptr && (MODE == 0 || !norm)) {
            if (p.dy_sum_acc) {
              float t[8];
              load8f(p.dy_sum, n, q.z, q.y, q.x, cg, t);
#pragma unroll
              for (int e = 0; e < 8; ++e) t[e] += g[e];
              store8f(p.dy_sum, n, q.z, q.y, q.x, cg, t);
            } else {
              store8f(p.dy_sum, n, q.z, q.y, q.x, cg, g);
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] *= oscale;
        float rr[8];
        if (rba) load8(p.res, n, q.z, q.y, q.x, cg, rr);
        // activation derivative
        if (norm) {
#pragma unroll
          for (int e = 0; e < 8; ++e) xv[e] = (xv[e] - mean[e]) * rstd[e];  // xhat
          if (p.act == GB_ACT_RELU || p.act == GB_ACT_LEAKY || p.act == GB_ACT_PRELU) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float pre = rba ? xv[e] + rr[e] : xv[e];
              if (!(pre > 0.f)) {
                if (MODE == 0 && p.act == GB_ACT_PRELU) sp[e] += g[e] * pre;
                g[e] *= (p.act == GB_ACT_RELU) ? 0.f : slope[e];
              }
            }
          }
        } else if (p.act != GB_ACT_NONE) {
          if (p.y.ptr != nullptr) {
            // activation only: derivative from the forward OUTPUT y (loaded into xv)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (p.act == GB_ACT_TANH) g[e] *= (1.f - xv[e] * xv[e]);
              else if (!(xv[e] > 0.f)) g[e] *= (p.act == GB_ACT_RELU) ? 0.f : slope[e];
            }
          } else {
            // derivative from the pre-activation x (+ res)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float pre = rba ? xv[e] + rr[e] : xv[e];
              if (p.act == GB_ACT_TANH) {
                const float th = tanhf(pre);
                g[e] *= (1.f - th * th);
              } else if (!(pre > 0.f)) {
                if (want_dprelu1) sp[e] += g[e] * pre;
                g[e] *= (p.act == GB_ACT_RELU) ? 0.f : slope[e];
              }
            }
          }
        }
        // residual added BEFORE the activation: its gradient is the masked g
        if (rba && p.dy_sum.ptr != nullptr && (MODE == 1 || !norm)) {
          if (p.dy_sum_acc) {
            float t[8];
            load8f(p.dy_sum, n, q.z, q.y, q.x, cg, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) t[e] += g[e];
            store8f(p.dy_sum, n, q.z, q.y, q.x, cg, t);
          } else {
            store8f(p.dy_sum, n, q.z, q.y, q.x, cg, g);
          }
        }
        if (MODE == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            s1[e] += g[e];
            s2[e] += g[e] * xv[e];
          }
        } else {
          float d[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) d[e] = norm ? rstd[e] * (g[e] - m1[e] - xv[e] * m2[e]) : g[e];
          if (p.dx_fp32_acc) {
            float t[8];
            load8f(p.dx, n, q.z, q.y, q.x, cg, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) t[e] += d[e];
            store8f(p.dx, n, q.z, q.y, q.x, cg, t);
            continue;
          }
          const uint4 packed = pack8(d);
          store8(p.dx, n, q.z, q.y, q.x, cg, packed);
          if (want_dbias) {  // bias gradient = sum over pixels of the fp32 dx (before the bf16 rounding of the
                             // stored operand: the sum of rounding errors is a random walk, not a gradient)
#pragma unroll
            for (int e = 0; e < 8; ++e) s1[e] += d[e];
          }
        }
      }
    }
  }
  if (MODE == 0) {
    if (slot < slots) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        red[(slot * x.C + cg * 8 + e) * 3 + 0] = s1[e];
        red[(slot * x.C + cg * 8 + e) * 3 + 1] = s2[e];
        red[(slot * x.C + cg * 8 + e) * 3 + 2] = sp[e];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < x.C; c += blockDim.x) {
      float a = 0.f, b = 0.f, d = 0.f;
      for (int k = 0; k < slots; ++k) {
        a += red[(k * x.C + c) * 3 + 0];
        b += red[(k * x.C + c) * 3 + 1];
        d += red[(k * x.C + c) * 3 + 2];
      }
      atomicAdd(p.bstats + ((int64_t)n * x.C + c) * 2 + 0, a);
      atomicAdd(p.bstats + ((int64_t)n * x.C + c) * 2 + 1, b);
      if (p.dprelu != nullptr) atomicAdd(p.dprelu + c, d);
    }
  }
  if (want_dbias || want_dprelu1) {
    if (slot < slots) {
#pragma unroll
      for (int e = 0; e < 8; ++e) red[slot * x.C + cg * 8 + e] = want_dbias ? s1[e] : sp[e];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < x.C; c += blockDim.x) {
      float a = 0.f;
      for (int k = 0; k < slots; ++k) a += red[k * x.C + c];
      atomicAdd((want_dbias ? p.dbias : p.dprelu) + c, a);
    }
  }
}

struct Launch {
  int threads, slots, ppb;
  dim3 grid;
};
Launch plan(const gb_view& x) {
  Launch L;
  const int C8 = x.C / 8;
  L.threads = 256;
  if (C8 > 256) L.threads = ((C8 + 31) / 32) * 32;
  L.slots = L.threads / C8;
  const int64_t P = (int64_t)x.D * x.H * x.W;
  // ONE resident wave: 148 SMs x 2 blocks (the kernels need ~128 registers). Each block streams its share of
  // the pixels with several loads in flight; more, shorter blocks only add per-block prologue / reduction tails
  // (measured: 3.5 waves of 10 us blocks = 38 us for a tensor that one wave streams in a quarter of that).
  int64_t blocks = (148 * 2 + x.N - 1) / x.N;
  if (blocks < 1) blocks = 1;
  int64_t ppb = (P + blocks - 1) / blocks;
  const int64_t quantum = 2 * L.slots;  // two pixels in flight per thread
  ppb = (ppb + quantum - 1) / quantum * quantum;
  L.ppb = (int)ppb;
  L.grid = dim3((unsigned)((P + ppb - 1) / ppb), x.N);
  return L;
}

bool same_extents(const gb_view& a, const gb_view& b) {
  return a.N == b.N && a.D == b.D && a.H == b.H && a.W == b.W && a.C == b.C;
}

}  // namespace

int gb_in_fwd_fast(const gb_in_fwd_params& p, cudaStream_t st);  // instnorm_fast.cu: -1 = not covered
int gb_in_bwd_fast(const gb_in_bwd_params& p, cudaStream_t st);
int gb_in_fwd_fast_v2(const gb_in_fwd_params& p, cudaStream_t st);  // instnorm_v2.cu: opt-in (knob 26)
int gb_in_bwd_onchip(const gb_in_bwd_params& p, cudaStream_t st);   // instnorm_v3.cu: opt-in (knob 24), -1 = not covered
int gb_in_bwd_fast_v2(const gb_in_bwd_params& p, cudaStream_t st);  // instnorm_v2.cu: opt-in (knob 22), -1 = not covered

extern "C" int gb_in_stats(const gb_view* x, float* stats, void* stream) {
  GB_CHECK(x && x->ptr && stats, "gb_in_stats: null pointer");
  GB_CHECK(x->C % 8 == 0 && x->C <= 2048, "gb_in_stats: bad channel count %d", x->C);
  Launch L = plan(*x);
  gb_klaunch(in_stats_kernel, L.grid, L.threads, sizeof(float) * 2 * L.slots * x->C, (cudaStream_t)stream, *x, stats, L.ppb);
  GB_LAUNCH_CHECK();
  return 0;
}

extern "C" int gb_in_fwd(const gb_in_fwd_params* p, void* stream) {
  GB_CHECK(p && p->x.ptr && p->y.ptr, "gb_in_fwd: null pointer");
  GB_CHECK(p->x.C % 8 == 0 && p->x.C <= 2048, "gb_in_fwd: bad channel count %d", p->x.C);
  GB_CHECK(same_extents(p->x, p->y), "gb_in_fwd: x/y extents differ");
  GB_CHECK(p->res.ptr == nullptr || same_extents(p->x, p->res), "gb_in_fwd: residual extents differ");
  GB_CHECK(p->act != GB_ACT_PRELU || p->prelu != nullptr, "gb_in_fwd: prelu slopes missing");
  GB_CHECK(p->y.pad == 0 || (p->y.H > p->y.pad && p->y.W > p->y.pad), "gb_in_fwd: reflection border larger than image");
  if (g_gb_knobs[26] != 0) {
    const int r = gb_in_fwd_fast_v2(*p, (cudaStream_t)stream);
    if (r >= 0) {
      ++g_gb_knobs[27];  // launches served by the second-generation forward (read back by the tests)
      return r;
    }
  }
  {
    const int r = gb_in_fwd_fast(*p, (cudaStream_t)stream);
    if (r >= 0) return r;
  }
  Launch L = plan(p->x);
  gb_klaunch(in_fwd_kernel, L.grid, L.threads, 0, (cudaStream_t)stream, *p, L.ppb);
  GB_LAUNCH_CHECK();
  return 0;
}

extern "C" int gb_in_bwd(const gb_in_bwd_params* p, void* stream) {
  GB_CHECK(p && p->x.ptr && p->dx.ptr, "gb_in_bwd: null pointer");
  GB_CHECK(p->dy_a.ptr || p->dy_b.ptr, "gb_in_bwd: no incoming gradient");
  GB_CHECK(p->x.C % 8 == 0 && p->x.C <= 2048, "gb_in_bwd: bad channel count %d", p->x.C);
  GB_CHECK(same_extents(p->x, p->dx), "gb_in_bwd: x/dx extents differ");
  GB_CHECK(p->dy_a.ptr == nullptr || same_extents(p->x, p->dy_a), "gb_in_bwd: dy_a extents differ");
  GB_CHECK(p->dy_b.ptr == nullptr || same_extents(p->x, p->dy_b), "gb_in_bwd: dy_b extents differ");
  GB_CHECK(p->stats == nullptr || p->bstats != nullptr, "gb_in_bwd: bstats workspace missing");
  GB_CHECK(!(p->dbias && p->dprelu && p->stats == nullptr), "gb_in_bwd: dbias and dprelu cannot both be reduced without a norm");
  GB_CHECK(!p->res_before_act || p->res.ptr == nullptr || same_extents(p->x, p->res), "gb_in_bwd: residual extents differ");
  cudaStream_t st = (cudaStream_t)stream;
  if (g_gb_knobs[24] != 0) {
    const int r = gb_in_bwd_onchip(*p, st);
    if (r >= 0) {
      ++g_gb_knobs[25];  // launches served by the on-chip kernel (read back by the tests)
      return r;
    }
  }
  if (g_gb_knobs[22] != 0) {
    const int r = gb_in_bwd_fast_v2(*p, st);
    if (r >= 0) {
      ++g_gb_knobs[23];  // launches served by the second-generation kernel (read back by the tests)
      return r;
    }
  }
  {
    const int r = gb_in_bwd_fast(*p, st);
    if (r >= 0) return r;
  }
  Launch L = plan(p->x);
  if (p->stats != nullptr) {
    gb_klaunch(in_bwd_kernel<0>, L.grid, L.threads, sizeof(float) * 3 * L.slots * p->x.C, st, *p, L.ppb);
    GB_LAUNCH_CHECK();
  }
  gb_klaunch(in_bwd_kernel<1>, L.grid, L.threads, (p->dbias || p->dprelu) ? sizeof(float) * L.slots * p->x.C : 0, st, *p, L.ppb);
  GB_LAUNCH_CHECK();
  return 0;
}
