// Fused PatchNCE loss (CUT): logits, diagonal mask, temperature, log-sum-exp and cross-entropy against class 0 in
// one kernel, and its backward in a second one.  Reference: ganslate/nn/losses/cut_losses.py:14-43
//   l_pos = q_r . k_r ;  l_neg[j] = q_r . k_(b,j)  (j == own patch -> -10) ;  out = cat(l_pos, l_neg) / T ;
//   loss_r = CE(out, 0) = logsumexp(out) - out[0]          (reduction 'none': one value per row)
// k is detached in the reference, so only dq is produced.  The problem is tiny (P = 256 patches, D = 256 features
// per layer) and latency bound; one block per row keeps it to a single launch per direction.
#include "gb_common.cuh"

namespace {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid = B*P rows, block = 256 threads. probs: [B*P][P+1] softmax of the logits (saved for backward).
__global__ void patchnce_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, int P, int D, float inv_T,
                                    float* __restrict__ loss, float* __restrict__ probs) {
  gb_pdl_enter();
  extern __shared__ float sh[];  // q row [D] | logits [P+1] | scratch [64]
  float* qs = sh;
  float* lg = sh + D;
  float* red = lg + P + 1;
  const int r = blockIdx.x;
  const int b = r / P, i = r - b * P;
  for (int d = threadIdx.x; d < D; d += blockDim.x) qs[d] = q[(int64_t)r * D + d];
  __syncthreads();
  for (int j = threadIdx.x; j <= P; j += blockDim.x) {
    // column 0 is the positive (own key), column 1 + jj the negatives of the same image
    const int jj = j - 1;
    float v;
    if (j > 0 && jj == i) {
      v = -10.0f;
    } else {
      const float* kr = k + (int64_t)(j == 0 ? r : b * P + jj) * D;
      float acc = 0.f;
      for (int d = 0; d < D; d += 4) {
        const float4 kk = __ldg(reinterpret_cast<const float4*>(kr + d));
        acc += qs[d] * kk.x + qs[d + 1] * kk.y + qs[d + 2] * kk.z + qs[d + 3] * kk.w;
      }
      v = acc;
    }
    lg[j] = v * inv_T;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int j = threadIdx.x; j <= P; j += blockDim.x) m = fmaxf(m, lg[j]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int j = threadIdx.x; j <= P; j += blockDim.x) s += __expf(lg[j] - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
  const float lse = m + __logf(s);
  if (threadIdx.x == 0) loss[r] = lse - lg[0];
  if (probs != nullptr)
    for (int j = threadIdx.x; j <= P; j += blockDim.x) probs[(int64_t)r * (P + 1) + j] = __expf(lg[j] - lse);
}

// dq_r = dloss_r / T * ( (p0 - 1) k_r + sum_{j != i} p_{1+j} k_(b,j) )
__global__ void patchnce_bwd_kernel(const float* __restrict__ k, const float* __restrict__ probs,
                                    const float* __restrict__ dloss, int P, int D, float inv_T, float* __restrict__ dq) {
  gb_pdl_enter();
  extern __shared__ float sh[];  // coefficients [P+1]
  const int r = blockIdx.x;
  const int b = r / P, i = r - b * P;
  const float g = dloss[r] * inv_T;
  for (int j = threadIdx.x; j <= P; j += blockDim.x) {
    float c = probs[(int64_t)r * (P + 1) + j];
    if (j == 0) c -= 1.f;
    if (j > 0 && j - 1 == i) c = 0.f;  // the masked diagonal entry is a constant
    sh[j] = c * g;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = sh[0] * __ldg(k + (int64_t)r * D + d);
    const float* kb = k + (int64_t)b * P * D + d;
    for (int j = 0; j < P; ++j) acc += sh[1 + j] * __ldg(kb + (int64_t)j * D);
    dq[(int64_t)r * D + d] = acc;
  }
}

}  // namespace

extern "C" int gb_patchnce_fwd(const float* q, const float* k, int B, int P, int D, float T, float* loss, float* probs,
                               void* stream) {
  GB_CHECK(q && k && loss, "gb_patchnce_fwd: null pointer");
  GB_CHECK(B > 0 && P > 0 && D > 0 && D % 4 == 0 && T > 0.f, "gb_patchnce_fwd: bad sizes B=%d P=%d D=%d", B, P, D);
  const size_t smem = sizeof(float) * (D + P + 1 + 64);
  GB_CHECK(smem <= 48 * 1024, "gb_patchnce_fwd: P + D too large");
  gb_klaunch(patchnce_fwd_kernel, B * P, 256, smem, (cudaStream_t)stream, q, k, P, D, 1.f / T, loss, probs);
  GB_LAUNCH_CHECK();
  return 0;
}

extern "C" int gb_patchnce_bwd(const float* k, const float* probs, const float* dloss, int B, int P, int D, float T,
                               float* dq, void* stream) {
  GB_CHECK(k && probs && dloss && dq, "gb_patchnce_bwd: null pointer");
  GB_CHECK(B > 0 && P > 0 && D > 0 && T > 0.f, "gb_patchnce_bwd: bad sizes");
  const size_t smem = sizeof(float) * (P + 1);
  GB_CHECK(smem <= 48 * 1024, "gb_patchnce_bwd: P too large");
  gb_klaunch(patchnce_bwd_kernel, B * P, 256, smem, (cudaStream_t)stream, k, probs, dloss, P, D, 1.f / T, dq);
  GB_LAUNCH_CHECK();
  return 0;
}
