// ReplicationPad3d forward / backward on channels-last views (the 3-D ResNet generators:
// ganslate/nn/generators/resnet/resnet3d.py:24,64,80,84 and piresnet3d.py:61,85,117).
//
// The 2-D generators use reflection padding, whose border the producing kernel writes itself (instnorm_fast.cu);
// a TMA box cannot clamp its coordinates, so replicate padding is materialised: one streaming copy forward
// (2 B read + 2 B written per padded element), one fold of the FP32 gradient backward.  HBM-bound byte work:
// 16-byte vectors, channel group fastest so that a warp touches consecutive addresses.
#include "gb_common.cuh"
#include "gb_geometry.h"

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// dst[n, z, y, x, :] = src[n, clamp(z - pz), clamp(y - py), clamp(x - px), :]
__global__ void __launch_bounds__(256) replicate_pad_fwd_kernel(gb_view src, gb_view dst, int pz, int py, int px) {
  gb_pdl_enter();
  const int C8 = dst.C >> 3;
  const int64_t total = (int64_t)dst.N * dst.D * dst.H * dst.W * C8;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(src.ptr);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst.ptr);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C8);
    int64_t pix = i / C8;
    const int x = (int)(pix % dst.W);
    pix /= dst.W;
    const int y = (int)(pix % dst.H);
    pix /= dst.H;
    const int z = (int)(pix % dst.D);
    const int n = (int)(pix / dst.D);
    const int sz = clampi(z - pz, 0, src.D - 1), sy = clampi(y - py, 0, src.H - 1), sx = clampi(x - px, 0, src.W - 1);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + gb_pix_offset(src, n, sz, sy, sx) + cg * 8));
    *reinterpret_cast<uint4*>(out + gb_pix_offset(dst, n, z, y, x) + cg * 8) = v;
  }
}

// range of padded coordinates that read source coordinate i of an axis of length n padded by p on both sides
__device__ __forceinline__ void readers(int i, int n, int p, int& lo, int& hi) {
  lo = (i == 0) ? 0 : i + p;
  hi = (i == n - 1) ? n - 1 + 2 * p : i + p;
}

// dsrc[n, z, y, x, :] += sum over padded positions (zz, yy, xx) that clamp onto (z, y, x) of ddst[n, zz, yy, xx, :]
__global__ void __launch_bounds__(256) replicate_pad_bwd_kernel(gb_view ddst, gb_view dsrc, int pz, int py, int px) {
  gb_pdl_enter();
  const int C4 = dsrc.C >> 2;
  const int64_t total = (int64_t)dsrc.N * dsrc.D * dsrc.H * dsrc.W * C4;
  const float* g = reinterpret_cast<const float*>(ddst.ptr);
  float* out = reinterpret_cast<float*>(dsrc.ptr);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C4);
    int64_t pix = i / C4;
    const int x = (int)(pix % dsrc.W);
    pix /= dsrc.W;
    const int y = (int)(pix % dsrc.H);
    pix /= dsrc.H;
    const int z = (int)(pix % dsrc.D);
    const int n = (int)(pix / dsrc.D);
    int z0, z1, y0, y1, x0, x1;
    readers(z, dsrc.D, pz, z0, z1);
    readers(y, dsrc.H, py, y0, y1);
    readers(x, dsrc.W, px, x0, x1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int zz = z0; zz <= z1; ++zz)
      for (int yy = y0; yy <= y1; ++yy)
        for (int xx = x0; xx <= x1; ++xx) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(g + gb_pix_offset(ddst, n, zz, yy, xx) + cg * 4));
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
    float4* o = reinterpret_cast<float4*>(out + gb_pix_offset(dsrc, n, z, y, x) + cg * 4);
    float4 prev = *o;
    prev.x += acc.x; prev.y += acc.y; prev.z += acc.z; prev.w += acc.w;
    *o = prev;
  }
}

int grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride loop: a multiple of the SM count, 16 blocks in flight per SM
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

bool vec_ok(const gb_view& v, int elem_bytes) {
  return ((uintptr_t)v.ptr % 16) == 0 && (v.sx * elem_bytes) % 16 == 0 && (v.sy * elem_bytes) % 16 == 0 &&
         (v.sz * elem_bytes) % 16 == 0 && (v.sn * elem_bytes) % 16 == 0;
}

}  // namespace

extern "C" int gb_replicate_pad_fwd(const gb_view* src, const gb_view* dst, int pz, int py, int px, void* stream) {
  GB_CHECK(src && dst && src->ptr && dst->ptr, "gb_replicate_pad_fwd: null pointer");
  GB_CHECK(pz >= 0 && py >= 0 && px >= 0, "gb_replicate_pad_fwd: negative padding");
  GB_CHECK(src->N == dst->N && src->C == dst->C && src->C % 8 == 0, "gb_replicate_pad_fwd: batch / channel mismatch");
  GB_CHECK(dst->D == src->D + 2 * pz && dst->H == src->H + 2 * py && dst->W == src->W + 2 * px,
           "gb_replicate_pad_fwd: dst extents (%d,%d,%d) are not src (%d,%d,%d) + 2 * pad", dst->D, dst->H, dst->W, src->D,
           src->H, src->W);
  GB_CHECK(vec_ok(*src, 2) && vec_ok(*dst, 2), "gb_replicate_pad_fwd: views must be 16-byte aligned");
  const int64_t total = (int64_t)dst->N * dst->D * dst->H * dst->W * (dst->C / 8);
  if (total == 0) return 0;
  gb_klaunch(replicate_pad_fwd_kernel, grid_for(total), 256, 0, (cudaStream_t)stream, *src, *dst, pz, py, px);
  GB_LAUNCH_CHECK();
  return 0;
}

extern "C" int gb_replicate_pad_bwd(const gb_view* ddst, const gb_view* dsrc, int pz, int py, int px, void* stream) {
  GB_CHECK(ddst && dsrc && ddst->ptr && dsrc->ptr, "gb_replicate_pad_bwd: null pointer");
  GB_CHECK(pz >= 0 && py >= 0 && px >= 0, "gb_replicate_pad_bwd: negative padding");
  GB_CHECK(dsrc->N == ddst->N && dsrc->C == ddst->C && dsrc->C % 8 == 0, "gb_replicate_pad_bwd: batch / channel mismatch");
  GB_CHECK(ddst->D == dsrc->D + 2 * pz && ddst->H == dsrc->H + 2 * py && ddst->W == dsrc->W + 2 * px,
           "gb_replicate_pad_bwd: gradient extents do not match");
  GB_CHECK(vec_ok(*dsrc, 4) && vec_ok(*ddst, 4), "gb_replicate_pad_bwd: views must be 16-byte aligned");
  const int64_t total = (int64_t)dsrc->N * dsrc->D * dsrc->H * dsrc->W * (dsrc->C / 4);
  if (total == 0) return 0;
  gb_klaunch(replicate_pad_bwd_kernel, grid_for(total), 256, 0, (cudaStream_t)stream, *ddst, *dsrc, pz, py, px);
  GB_LAUNCH_CHECK();
  return 0;
}
