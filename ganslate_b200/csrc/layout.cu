// Layout conversion at the network boundary: NC(D)HW fp32 (what ganslate's set_input hands to the networks,
// ganslate/nn/gans/unpaired/cyclegan.py:89-90) <-> channels-last bf16 views used by every kernel here.
#include "gb_common.cuh"
#include "gb_geometry.h"

namespace {

__device__ __forceinline__ void reflect_targets(const gb_view& v, int y, int x, int (&ys)[3], int (&xs)[3], int& ny,
                                                int& nx) {
  const int p = v.pad;
  ny = nx = 1;
  ys[0] = y;
  xs[0] = x;
  if (p > 0) {
    if (y >= 1 && y <= p) ys[ny++] = -y;
    if (y <= v.H - 2 && y >= v.H - 1 - p) ys[ny++] = 2 * (v.H - 1) - y;
    if (x >= 1 && x <= p) xs[nx++] = -x;
    if (x <= v.W - 2 && x >= v.W - 1 - p) xs[nx++] = 2 * (v.W - 1) - x;
  }
}

__global__ void nchw_to_cl_kernel(const float* __restrict__ src, int C, gb_view dst, gb_view pre, int dst_fp32) {
  gb_pdl_enter();
  const int64_t P = (int64_t)dst.D * dst.H * dst.W;
  const int64_t total = (int64_t)dst.N * P;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst.ptr);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / P);
    int64_t pix = i - (int64_t)n * P;
    const int x = (int)(pix % dst.W);
    const int y = (int)((pix / dst.W) % dst.H);
    const int z = (int)(pix / ((int64_t)dst.W * dst.H));
    int ys[3], xs[3], ny, nx;
    reflect_targets(dst, y, x, ys, xs, ny, nx);
    for (int cg = 0; cg < dst.C / 8; ++cg) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = cg * 8 + e;
        f[e] = c < C ? __ldg(src + ((int64_t)n * C + c) * P + pix) : 0.f;
      }
      if (pre.ptr != nullptr) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(pre.ptr) +
                                                             gb_pix_offset(pre, n, z, y, x) + cg * 8));
        float2 t;
        float th;
        t = unpack_bf16x2(u.x); th = tanhf(t.x); f[0] *= 1.f - th * th; th = tanhf(t.y); f[1] *= 1.f - th * th;
        t = unpack_bf16x2(u.y); th = tanhf(t.x); f[2] *= 1.f - th * th; th = tanhf(t.y); f[3] *= 1.f - th * th;
        t = unpack_bf16x2(u.z); th = tanhf(t.x); f[4] *= 1.f - th * th; th = tanhf(t.y); f[5] *= 1.f - th * th;
        t = unpack_bf16x2(u.w); th = tanhf(t.x); f[6] *= 1.f - th * th; th = tanhf(t.y); f[7] *= 1.f - th * th;
      }
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]);
      o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]);
      o.w = pack_bf16x2(f[6], f[7]);
      for (int a = 0; a < ny; ++a)
        for (int b = 0; b < nx; ++b) {
          if (dst_fp32) {
            float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(dst.ptr) +
                                                   gb_pix_offset(dst, n, z, ys[a], xs[b]) + cg * 8);
            o4[0] = make_float4(f[0], f[1], f[2], f[3]);
            o4[1] = make_float4(f[4], f[5], f[6], f[7]);
          } else {
            *reinterpret_cast<uint4*>(out + gb_pix_offset(dst, n, z, ys[a], xs[b]) + cg * 8) = o;
          }
        }
    }
  }
}

__global__ void cl_to_nchw_kernel(gb_view src, float* __restrict__ dst, int C, int fold, int act, int src_fp32) {
  gb_pdl_enter();
  const int64_t P = (int64_t)src.D * src.H * src.W;
  const int64_t total = (int64_t)src.N * P;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(src.ptr);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / P);
    int64_t pix = i - (int64_t)n * P;
    const int x = (int)(pix % src.W);
    const int y = (int)((pix / src.W) % src.H);
    const int z = (int)(pix / ((int64_t)src.W * src.H));
    int ys[3], xs[3], ny = 1, nx = 1;
    ys[0] = y;
    xs[0] = x;
    if (fold) reflect_targets(src, y, x, ys, xs, ny, nx);
    for (int cg = 0; cg * 8 < C; ++cg) {
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int a = 0; a < ny; ++a)
        for (int b = 0; b < nx; ++b) {
          if (src_fp32) {
            const float4* q4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src.ptr) +
                                                               gb_pix_offset(src, n, z, ys[a], xs[b]) + cg * 8);
            const float4 u0 = __ldg(q4), u1 = __ldg(q4 + 1);
            f[0] += u0.x; f[1] += u0.y; f[2] += u0.z; f[3] += u0.w;
            f[4] += u1.x; f[5] += u1.y; f[6] += u1.z; f[7] += u1.w;
            continue;
          }
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + gb_pix_offset(src, n, z, ys[a], xs[b]) + cg * 8));
          float2 t;
          t = unpack_bf16x2(u.x); f[0] += t.x; f[1] += t.y;
          t = unpack_bf16x2(u.y); f[2] += t.x; f[3] += t.y;
          t = unpack_bf16x2(u.z); f[4] += t.x; f[5] += t.y;
          t = unpack_bf16x2(u.w); f[6] += t.x; f[7] += t.y;
        }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = cg * 8 + e;
        if (c < C) dst[((int64_t)n * C + c) * P + pix] = (act == GB_ACT_TANH) ? tanhf(f[e]) : f[e];
      }
    }
  }
}

}  // namespace

extern "C" int gb_nchw_to_cl(const float* src, int C, const gb_view* dst, const gb_view* pre, int dst_fp32,
                             void* stream) {
  GB_CHECK(src && dst && dst->ptr, "gb_nchw_to_cl: null pointer");
  GB_CHECK(dst->C % 8 == 0 && C <= dst->C && C >= 1, "gb_nchw_to_cl: bad channel counts %d -> %d", C, dst->C);
  GB_CHECK(dst->pad == 0 || (dst->H > dst->pad && dst->W > dst->pad), "gb_nchw_to_cl: border larger than image");
  const int64_t total = (int64_t)dst->N * dst->D * dst->H * dst->W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  gb_view pv = {};
  if (pre != nullptr) pv = *pre;
  gb_klaunch(nchw_to_cl_kernel, blocks, 256, 0, (cudaStream_t)stream, src, C, *dst, pv, dst_fp32);
  GB_LAUNCH_CHECK();
  return 0;
}

extern "C" int gb_cl_to_nchw(const gb_view* src, float* dst, int C, int fold, int act, int src_fp32, void* stream) {
  GB_CHECK(src && src->ptr && dst, "gb_cl_to_nchw: null pointer");
  GB_CHECK(src->C % 8 == 0 && C <= src->C && C >= 1, "gb_cl_to_nchw: bad channel counts %d <- %d", C, src->C);
  const int64_t total = (int64_t)src->N * src->D * src->H * src->W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  gb_klaunch(cl_to_nchw_kernel, blocks, 256, 0, (cudaStream_t)stream, *src, dst, C, fold, act, src_fp32);
  GB_LAUNCH_CHECK();
  return 0;
}
