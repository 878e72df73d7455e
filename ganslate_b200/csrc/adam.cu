// Multi-tensor Adam step (torch.optim.Adam semantics: no weight decay, no amsgrad), fp32 master weights.
//   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Replaces the optimizer.step() calls of the recipes (ganslate/nn/gans/unpaired/cyclegan.py:76-82,108,121): the
// reference's torch.optim.Adam walks the parameter list with several elementwise passes; here every parameter of
// an optimizer is updated by one launch per GB_ADAM_BATCH tensors, HBM-bound at 28 bytes per element
// (read p, g, m, v; write p, m, v).  lr and the step count are read from device memory so that the launch can be
// replayed inside a CUDA graph while the scheduler changes lr.
#include "gb_common.cuh"

namespace {

__global__ void __launch_bounds__(256) adam_multi_kernel(const __grid_constant__ gb_adam_batch b) {
  gb_pdl_enter();
  const gb_adam_item& it = b.item[blockIdx.y];
  const float lr = *b.lr;
  const float t = *b.step;  // already incremented for this step
  const float bc1 = 1.f - powf(b.beta1, t);
  const float bc2 = 1.f - powf(b.beta2, t);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float b1 = b.beta1, b2 = b.beta2, eps = b.eps;
  const int64_t n = it.n;
  const int64_t n4 = it.vec4 ? (n >> 2) : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(it.p);
  float4* m4 = reinterpret_cast<float4*>(it.m);
  float4* v4 = reinterpret_cast<float4*>(it.v);
  const float4* g4 = reinterpret_cast<const float4*>(it.g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = p4[i], m = m4[i], v = v4[i];
    const float4 g = g4[i];
    float* pp = &p.x; float* mm = &m.x; float* vv = &v.x;
    const float* gg = &g.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mm[e] = b1 * mm[e] + (1.f - b1) * gg[e];
      vv[e] = b2 * vv[e] + (1.f - b2) * gg[e] * gg[e];
      pp[e] -= step_size * mm[e] / (sqrtf(vv[e]) * inv_sqrt_bc2 + eps);
    }
    p4[i] = p;
    m4[i] = m;
    v4[i] = v;
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float g = it.g[i];
    const float m = b1 * it.m[i] + (1.f - b1) * g;
    const float v = b2 * it.v[i] + (1.f - b2) * g * g;
    it.m[i] = m;
    it.v[i] = v;
    it.p[i] -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
  }
}

}  // namespace

extern "C" int gb_adam_multi(const gb_adam_batch* b, void* stream) {
  GB_CHECK(b && b->count >= 1 && b->count <= GB_ADAM_BATCH, "gb_adam_multi: bad item count");
  GB_CHECK(b->lr && b->step, "gb_adam_multi: lr / step must be device pointers");
  int64_t mx = 0;
  for (int i = 0; i < b->count; ++i) {
    const gb_adam_item& it = b->item[i];
    GB_CHECK(it.p && it.g && it.m && it.v && it.n >= 0, "gb_adam_multi: null pointer in item %d", i);
    GB_CHECK(!it.vec4 || ((((uintptr_t)it.p | (uintptr_t)it.g | (uintptr_t)it.m | (uintptr_t)it.v) & 15) == 0),
             "gb_adam_multi: item %d flagged vec4 but not 16-byte aligned", i);
    mx = it.n > mx ? it.n : mx;
  }
  int blocks = (int)((mx / 4 + 255) / 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  gb_klaunch(adam_multi_kernel, dim3(blocks, b->count), 256, 0, (cudaStream_t)stream, *b);
  GB_LAUNCH_CHECK();
  return 0;
}
