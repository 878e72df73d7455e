// Streaming fast paths of the fused InstanceNorm kernels (instnorm.cu keeps the fully general versions).
//
// Covered here: the patterns of the 2-D / 3-D generators and discriminators -- normalise + ReLU / LeakyReLU / none
// (+ residual added after the activation, + reflection border of the result) forward, and the matching backward
// with the gradient arriving on the (possibly reflection-padded) output domain, an optional residual-gradient
// accumulation and the bias-gradient column sum.  PReLU, residual-before-activation, scaled outputs and fp32
// destinations stay on the general kernels.
//
// Design: HBM-bound streaming.  A thread owns 4 consecutive channels (8 B of bf16, 16 B of fp32) so the per-channel
// constants cost 16 registers instead of 32; FU pixels per thread are loaded back to back before any arithmetic
// (all loads independent), three 256-thread blocks per SM are resident, and the grid is one resident wave.  The
// backward reduction and apply passes run in ONE launch when the grid is co-resident: blocks publish their
// partial sums with atomics, meet at a grid barrier (cooperative launch guarantees co-residency) and re-read the
// tensors, which at the sizes where launch latency matters are still in L2.
#include "gb_common.cuh"
#include "gb_geometry.h"

namespace {

constexpr int FU = 8;        // pixels in flight per thread (forward: 8 B + 8 B per pixel)
constexpr int BU = 4;        // backward: 16 B + 8 B (+ 16 B) per pixel
constexpr int FTHREADS = 256;
constexpr int FBLOCKS_PER_SM = 3;

struct FastGeom {
  gb_fastdiv divW;           // pixel -> (row, x) for the 2-D (bordered) addressing mode
  int ppb;                   // pixels per block
  int nblocks;               // blocks per image
  int total_blocks;          // grid size (fused mode: barrier target)
};

__device__ __forceinline__ uint2 ld4_bf16(const void* base, int64_t off) {
  return __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + off));
}
__device__ __forceinline__ void unpack4(const uint2& u, float (&f)[4]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
}
// coherent (L2) variants: used for data another block of the same launch may have written before the grid barrier
__device__ __forceinline__ float4 ld4_f32_cg(const void* base, int64_t off) {
  return __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off));
}
__device__ __forceinline__ void st4_f32(void* base, int64_t off, const float (&f)[4]) {
  *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off) = make_float4(f[0], f[1], f[2], f[3]);
}
__device__ __forceinline__ uint2 pack4(const float (&f)[4]) {
  uint2 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  return o;
}
__device__ __forceinline__ void st4_bf16(void* base, int64_t off, const uint2& o) {
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = o;
}

// element offset of pixel `pix` (index inside image n, row-major over the interior) in view v.
// LINEAR: every view is pixel-linear (no border, full rows): offset = n*sn + pix*sx.
template <bool LINEAR>
__device__ __forceinline__ int64_t pix_off(const gb_view& v, int n, uint32_t pix, int y, int x) {
  if (LINEAR) return (int64_t)n * v.sn + (int64_t)pix * v.sx;
  return (int64_t)n * v.sn + (int64_t)y * v.sy + (int64_t)x * v.sx;
}

// gradient on a reflection-padded domain folded onto interior pixel (y, x): adds the border positions that mirror
// onto it (the centre value is loaded by the caller)
__device__ __forceinline__ void add_mirrors4(const gb_view& v, int n, int y, int x, int c, float (&f)[4]) {
  const int p = v.pad;
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
      if (a == 0 && b == 0) continue;
      const float4 t = ld4_f32_cg(v.ptr, (int64_t)n * v.sn + (int64_t)ys[a] * v.sy + (int64_t)xs[b] * v.sx + c);
      f[0] += t.x; f[1] += t.y; f[2] += t.z; f[3] += t.w;
    }
}

// store to (y, x) and to every border position that reflects onto it
__device__ __forceinline__ void st4_reflect(const gb_view& v, int n, int y, int x, int c, const uint2& o) {
  const int p = v.pad;
  int ys[3], xs[3], ny = 1, nx = 1;
  ys[0] = y;
  xs[0] = x;
  if (y >= 1 && y <= p) ys[ny++] = -y;
  if (y <= v.H - 2 && y >= v.H - 1 - p) ys[ny++] = 2 * (v.H - 1) - y;
  if (x >= 1 && x <= p) xs[nx++] = -x;
  if (x <= v.W - 2 && x >= v.W - 1 - p) xs[nx++] = 2 * (v.W - 1) - x;
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) st4_bf16(v.ptr, (int64_t)n * v.sn + (int64_t)ys[a] * v.sy + (int64_t)xs[b] * v.sx + c, o);
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
template <bool RES, bool LINEAR>
__global__ void __launch_bounds__(FTHREADS, LINEAR ? FBLOCKS_PER_SM : 2)
in_fwd_fast_kernel(const __grid_constant__ gb_in_fwd_params p, const __grid_constant__ FastGeom g, float neg_slope) {
  const gb_view& x = p.x;
  const int C4 = x.C >> 2;
  const int slots = FTHREADS / C4;
  const int cg = threadIdx.x % C4;
  const int slot = threadIdx.x / C4;
  if (slot >= slots) return;
  const int c = cg * 4;
  const int n = blockIdx.y;
  const uint32_t P = (uint32_t)x.D * x.H * x.W;
  const uint32_t p0 = blockIdx.x * (uint32_t)g.ppb;
  const uint32_t p1 = min(P, p0 + (uint32_t)g.ppb);
  float mean[4], rstd[4];
  const float invP = 1.f / (float)P;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    mean[e] = 0.f;
    rstd[e] = 1.f;
  }
  if (p.stats != nullptr) {
    const float4 a = *reinterpret_cast<const float4*>(p.stats + ((int64_t)n * x.C + c) * 2);
    const float4 b = *reinterpret_cast<const float4*>(p.stats + ((int64_t)n * x.C + c) * 2 + 4);
    const float s[4] = {a.x, a.z, b.x, b.z}, ss[4] = {a.y, a.w, b.y, b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float m = s[e] * invP;
      mean[e] = m;
      rstd[e] = rsqrtf(fmaxf(ss[e] * invP - m * m, 0.f) + p.eps);
    }
  }
  for (uint32_t base = p0 + slot; base < p1; base += (uint32_t)(FU * slots)) {
    uint2 fx[FU], fr[FU];
    int yy[FU], xx[FU];
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      const uint32_t pix = base + (uint32_t)(u * slots);
      if (pix < p1) {
        if (!LINEAR) {
          const uint32_t row = gb_div(pix, g.divW);
          yy[u] = (int)row;
          xx[u] = (int)(pix - row * g.divW.d);
        } else {
          yy[u] = xx[u] = 0;
        }
        fx[u] = ld4_bf16(x.ptr, pix_off<LINEAR>(x, n, pix, yy[u], xx[u]) + c);
        if (RES) fr[u] = ld4_bf16(p.res.ptr, pix_off<LINEAR>(p.res, n, pix, yy[u], xx[u]) + c);
      }
    }
#pragma unroll
    for (int u = 0; u < FU; ++u) {
      const uint32_t pix = base + (uint32_t)(u * slots);
      if (pix < p1) {
        float f[4], r[4];
        unpack4(fx[u], f);
        if (RES) unpack4(fr[u], r);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = (f[e] - mean[e]) * rstd[e];
          v = v > 0.f ? v : v * neg_slope;
          if (RES) v += r[e];
          f[e] = v;
        }
        const uint2 o = pack4(f);
        if (!LINEAR && p.y.pad > 0) st4_reflect(p.y, n, yy[u], xx[u], c, o);
        else st4_bf16(p.y.ptr, pix_off<LINEAR>(p.y, n, pix, yy[u], xx[u]) + c, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// g = dy (folded from the padded domain when dy.pad > 0);  residual gradient: dy_sum += g (unmasked);
// g *= (xhat > 0 ? 1 : neg_slope);  pass 0: bstats += (sum g, sum g*xhat);
// pass 1: dx = rstd * (g - mean(g) - xhat * mean(g*xhat)) -> bf16, dbias += column sums of the fp32 dx.
// PASS 0 / 1 = the two passes as separate launches, PASS 2 = both in one launch around a grid barrier.
template <bool RES, bool LINEAR>
__device__ __forceinline__ void in_bwd_fast_pass(const gb_in_bwd_params& p, const FastGeom& g, float neg_slope, int pass,
                                                 bool first_pass_of_fused, float* red) {
  const gb_view& x = p.x;
  const gb_view& dy = p.dy_b;
  const int C4 = x.C >> 2;
  const int slots = FTHREADS / C4;
  const int cg = threadIdx.x % C4;
  const int slot = threadIdx.x / C4;
  const int c = cg * 4;
  const int n = blockIdx.y;
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
  if (slot < slots) {
    for (uint32_t base = p0 + slot; base < p1; base += (uint32_t)(BU * slots)) {
      float4 lg[BU], lr[BU];
      uint2 lx[BU];
      int yy[BU], xx[BU];
#pragma unroll
      for (int u = 0; u < BU; ++u) {
        const uint32_t pix = base + (uint32_t)(u * slots);
        if (pix < p1) {
          if (!LINEAR) {
            const uint32_t row = gb_div(pix, g.divW);
            yy[u] = (int)row;
            xx[u] = (int)(pix - row * g.divW.d);
          } else {
            yy[u] = xx[u] = 0;
          }
          lg[u] = ld4_f32_cg(dy.ptr, pix_off<LINEAR>(dy, n, pix, yy[u], xx[u]) + c);
          lx[u] = ld4_bf16(x.ptr, pix_off<LINEAR>(x, n, pix, yy[u], xx[u]) + c);
          if (do_res) lr[u] = ld4_f32_cg(p.dy_sum.ptr, pix_off<LINEAR>(p.dy_sum, n, pix, yy[u], xx[u]) + c);
        }
      }
#pragma unroll
      for (int u = 0; u < BU; ++u) {
        const uint32_t pix = base + (uint32_t)(u * slots);
        if (pix < p1) {
          float gg[4] = {lg[u].x, lg[u].y, lg[u].z, lg[u].w}, xv[4];
          unpack4(lx[u], xv);
          if (!LINEAR && dy.pad > 0) add_mirrors4(dy, n, yy[u], xx[u], c, gg);
          if (do_res) {
            const float rs[4] = {lr[u].x + gg[0], lr[u].y + gg[1], lr[u].z + gg[2], lr[u].w + gg[3]};
            st4_f32(p.dy_sum.ptr, pix_off<LINEAR>(p.dy_sum, n, pix, yy[u], xx[u]) + c, rs);
          }
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
            const uint2 o = pack4(d);
            st4_bf16(p.dx.ptr, pix_off<LINEAR>(p.dx, n, pix, yy[u], xx[u]) + c, o);
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

template <bool RES, bool LINEAR, int PASS>
__global__ void __launch_bounds__(FTHREADS, LINEAR ? FBLOCKS_PER_SM : 2)
in_bwd_fast_kernel(const __grid_constant__ gb_in_bwd_params p, const __grid_constant__ FastGeom g, float neg_slope) {
  extern __shared__ float red[];  // [slots][C][2]
  if (PASS == 2) {
    in_bwd_fast_pass<RES, LINEAR>(p, g, neg_slope, 0, true, red);
    unsigned int* counter = reinterpret_cast<unsigned int*>(p.bstats + (int64_t)p.x.N * p.x.C * 2);
    grid_barrier(counter, (unsigned int)g.total_blocks);
    in_bwd_fast_pass<RES, LINEAR>(p, g, neg_slope, 1, false, red);
  } else {
    in_bwd_fast_pass<RES, LINEAR>(p, g, neg_slope, PASS, true, red);
  }
}

// ------------------------------------------------------------------------------------------------ host side
bool pixel_linear(const gb_view& v) {
  return v.pad == 0 && v.sy == (int64_t)v.W * v.sx && (v.D == 1 || v.sz == (int64_t)v.H * v.sy);
}
bool aligned(const gb_view& v, int elem_bytes) {
  // 4-channel vectors: 8 B (bf16) / 16 B (fp32) alignment of every pixel's channel-slice start
  const int64_t a = (elem_bytes == 2) ? 4 : 4;
  return ((uintptr_t)v.ptr % (elem_bytes * 4)) == 0 && v.sx % a == 0 && v.sy % a == 0 && v.sz % a == 0 && v.sn % a == 0;
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
FastGeom plan(const gb_view& x, int cap, bool* fits) {
  FastGeom g;
  const int C4 = x.C / 4;
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
  g.divW = gb_make_fastdiv((uint32_t)x.W);
  return g;
}

template <bool RES, bool LINEAR>
int launch_fwd(const gb_in_fwd_params& p, float neg_slope, cudaStream_t st) {
  bool fits;
  const FastGeom g = plan(p.x, num_sms() * (LINEAR ? FBLOCKS_PER_SM : 2), &fits);
  in_fwd_fast_kernel<RES, LINEAR><<<dim3(g.nblocks, p.x.N), FTHREADS, 0, st>>>(p, g, neg_slope);
  GB_LAUNCH_CHECK();
  return 0;
}

template <bool RES, bool LINEAR>
int launch_bwd(const gb_in_bwd_params& p, float neg_slope, cudaStream_t st) {
  const int C4 = p.x.C / 4;
  const int slots = FTHREADS / C4;
  const size_t smem = sizeof(float) * 2 * slots * p.x.C;
  static int occ = -1;  // co-resident blocks per SM of the single-launch kernel (per instantiation)
  static size_t occ_smem = 0;
  if (occ < 0 || occ_smem != smem) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, in_bwd_fast_kernel<RES, LINEAR, 2>, FTHREADS, smem) != cudaSuccess) o = 0;
    cudaGetLastError();
    occ = o;
    occ_smem = smem;
  }
  bool fits = false;
  FastGeom g = plan(p.x, num_sms() * (occ > 0 ? occ : (LINEAR ? FBLOCKS_PER_SM : 2)), &fits);
  const dim3 grid(g.nblocks, p.x.N);
  if (occ > 0 && fits && g_gb_knobs[6] == 0) {
    void* args[] = {(void*)&p, (void*)&g, (void*)&neg_slope};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)in_bwd_fast_kernel<RES, LINEAR, 2>, grid, dim3(FTHREADS), args,
                                                smem, st);
    if (e == cudaSuccess) {
      __atomic_fetch_add(&g_gb_launches, 1ull, __ATOMIC_RELAXED);
      return 0;
    }
    cudaGetLastError();  // cooperative launch not possible here: fall through to two launches
  }
  in_bwd_fast_kernel<RES, LINEAR, 0><<<grid, FTHREADS, smem, st>>>(p, g, neg_slope);
  GB_LAUNCH_CHECK();
  in_bwd_fast_kernel<RES, LINEAR, 1><<<grid, FTHREADS, smem, st>>>(p, g, neg_slope);
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
  if (x.C % 4 != 0 || x.C / 4 > FTHREADS || (int64_t)x.D * x.H * x.W >= (1ll << 31)) return -1;
  const bool has_res = p.res.ptr != nullptr;
  if (!aligned(x, 2) || !aligned(p.y, 2) || (has_res && !aligned(p.res, 2))) return -1;
  if (p.stats != nullptr && ((uintptr_t)p.stats % 16 != 0 || (x.C * 2) % 4 != 0)) return -1;
  const bool linear = pixel_linear(x) && pixel_linear(p.y) && (!has_res || pixel_linear(p.res));
  if (!linear && x.D != 1) return -1;  // bordered addressing is 2-D
  if (linear) return has_res ? launch_fwd<true, true>(p, ns, st) : launch_fwd<false, true>(p, ns, st);
  return has_res ? launch_fwd<true, false>(p, ns, st) : launch_fwd<false, false>(p, ns, st);
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
  if (!aligned(x, 2) || !aligned(p.dx, 2) || !aligned(p.dy_b, 4) || (has_res && !aligned(p.dy_sum, 4))) return -1;
  if ((uintptr_t)p.stats % 16 != 0 || (uintptr_t)p.bstats % 16 != 0) return -1;
  const bool linear = pixel_linear(x) && pixel_linear(p.dx) && pixel_linear(p.dy_b) && (!has_res || pixel_linear(p.dy_sum));
  if (!linear && x.D != 1) return -1;
  if (linear) return has_res ? launch_bwd<true, true>(p, ns, st) : launch_bwd<false, true>(p, ns, st);
  return has_res ? launch_bwd<true, false>(p, ns, st) : launch_bwd<false, false>(p, ns, st);
}
