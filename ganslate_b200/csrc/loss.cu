// Fused loss reductions: value and gradient in one pass over the operands.
//   LSGAN  mean((p - t)^2), t constant        -- ganslate/nn/losses/adversarial_loss.py:29,60-62
//   L1     mean(|a - b|)                      -- ganslate/nn/losses/cyclegan_losses.py:64,73,97; pix2pix_losses.py:15
#include "gb_common.cuh"

namespace {

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sh[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;
}

__global__ void mse_const_kernel(const float* __restrict__ pred, float target, int64_t n, float inv_n,
                                 float* __restrict__ loss, float* __restrict__ grad) {
  gb_pdl_enter();
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = pred[i] - target;
    acc += d * d;
    if (grad) grad[i] = 2.f * d * inv_n;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_n);
}

__global__ void l1_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float inv_n,
                          float* __restrict__ loss, float* __restrict__ grad) {
  gb_pdl_enter();
  float acc = 0.f;
  const int64_t n4 = n >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  float4* g4 = reinterpret_cast<float4*>(grad);
  const bool vec = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)grad) & 15) == 0;
  if (vec) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
      const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
      acc += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
      if (grad) {
        float4 g;
        g.x = d0 > 0.f ? inv_n : (d0 < 0.f ? -inv_n : 0.f);
        g.y = d1 > 0.f ? inv_n : (d1 < 0.f ? -inv_n : 0.f);
        g.z = d2 > 0.f ? inv_n : (d2 < 0.f ? -inv_n : 0.f);
        g.w = d3 > 0.f ? inv_n : (d3 < 0.f ? -inv_n : 0.f);
        g4[i] = g;
      }
    }
  }
  const int64_t tail0 = vec ? (n4 << 2) : 0;
  for (int64_t i = tail0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    acc += fabsf(d);
    if (grad) grad[i] = d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_n);
}

int grid_for(int64_t n, int per_thread) {
  int64_t b = (n + 256ll * per_thread - 1) / (256ll * per_thread);
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int gb_mse_const(const float* pred, float target, int64_t n, float* loss, float* grad, void* stream) {
  GB_CHECK(pred && loss && n > 0, "gb_mse_const: bad arguments");
  gb_klaunch(mse_const_kernel, grid_for(n, 4), 256, 0, (cudaStream_t)stream, pred, target, n, 1.f / (float)n, loss, grad);
  GB_LAUNCH_CHECK();
  return 0;
}

extern "C" int gb_l1(const float* a, const float* b, int64_t n, float* loss, float* grad_a, void* stream) {
  GB_CHECK(a && b && loss && n > 0, "gb_l1: bad arguments");
  gb_klaunch(l1_kernel, grid_for(n, 16), 256, 0, (cudaStream_t)stream, a, b, n, 1.f / (float)n, loss, grad_a);
  GB_LAUNCH_CHECK();
  return 0;
}
