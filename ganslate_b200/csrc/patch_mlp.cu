// CUT's FeaturePatchMLP as fused kernels (ganslate/nn/gans/unpaired/cut.py:229-294): gather `P` positions of a
// feature map (the same ids for every image), Linear(C -> nc) + ReLU + Linear(nc -> nc), L2-normalise each row.
//   x[r][c]  = feat[n][c][id[p]]                      r = n * P + p
//   h        = relu(x W1^T + b1)                      W1: (nc, C)  torch.nn.Linear layout
//   z        = h W2^T + b2                            W2: (nc, nc)
//   y        = z / (||z||_2 + 1e-7)
// The problem is tiny (P = 256 rows per image, C <= 256, nc = 256: 33 MFLOP per feature and batch item) and sits
// between two tensor-core passes; fp32 CUDA-core FMAs keep it bit-comparable with the fp32 reference (no TF32, no
// library GEMM) and the whole forward is ONE launch per feature, the backward two.  Launch-latency bound by design.
#include "gb_common.cuh"

namespace {

constexpr int ROWS = 8;       // rows of x per CTA (every CTA streams W1 and W2 once from L2)
constexpr int THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

// grid = ceil(R / ROWS); smem: xs[ROWS][C] | hs[ROWS][nc] | zs[ROWS][nc] | red[32]
__global__ void __launch_bounds__(THREADS) patch_mlp_fwd_kernel(
    const float* __restrict__ feat, const int64_t* __restrict__ ids, int N, int C, int64_t F, int P,
    const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
    const float* __restrict__ b2, int nc, float* __restrict__ xg, float* __restrict__ h, float* __restrict__ z,
    float* __restrict__ y) {
  gb_pdl_enter();
  extern __shared__ float sh[];
  float* xs = sh;
  float* hs = xs + ROWS * C;
  float* zs = hs + ROWS * nc;
  float* red = zs + ROWS * nc;
  const int R = N * P;
  const int r0 = blockIdx.x * ROWS;
  const int tid = threadIdx.x;
  for (int i = tid; i < ROWS * C; i += THREADS) {
    const int rr = i / C, c = i - rr * C;
    const int r = r0 + rr;
    float v = 0.f;
    if (r < R) {
      const int n = r / P, p = r - n * P;
      v = feat[((int64_t)n * C + c) * F + ids[p]];
      xg[(int64_t)r * C + c] = v;
    }
    xs[i] = v;
  }
  __syncthreads();
  for (int j = tid; j < nc; j += THREADS) {
    float acc[ROWS];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) acc[rr] = b1[j];
    const float* w = W1 + (int64_t)j * C;
    for (int c = 0; c < C; ++c) {
      const float wv = __ldg(w + c);
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) acc[rr] = fmaf(xs[rr * C + c], wv, acc[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
      const float v = fmaxf(acc[rr], 0.f);
      hs[rr * nc + j] = v;
      if (r0 + rr < R) h[(int64_t)(r0 + rr) * nc + j] = v;
    }
  }
  __syncthreads();
  for (int j = tid; j < nc; j += THREADS) {
    float acc[ROWS];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) acc[rr] = b2[j];
    const float* w = W2 + (int64_t)j * nc;
    for (int k = 0; k < nc; ++k) {
      const float wv = __ldg(w + k);
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) acc[rr] = fmaf(hs[rr * nc + k], wv, acc[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
      zs[rr * nc + j] = acc[rr];
      if (r0 + rr < R) z[(int64_t)(r0 + rr) * nc + j] = acc[rr];
    }
  }
  __syncthreads();
  for (int rr = 0; rr < ROWS; ++rr) {
    float s = 0.f;
    for (int j = tid; j < nc; j += THREADS) s += zs[rr * nc + j] * zs[rr * nc + j];
    s = block_sum(s, red);
    const float inv = 1.f / (sqrtf(s) + 1e-7f);
    if (r0 + rr < R)
      for (int j = tid; j < nc; j += THREADS) y[(int64_t)(r0 + rr) * nc + j] = zs[rr * nc + j] * inv;
  }
}

// Row-wise part of the backward: dz (L2-norm), dh (ReLU mask), dx scattered into dfeat (zero-initialised by the
// caller; ids are distinct within a feature, so no two rows of one image meet).  grid = ceil(R / ROWS).
// smem: dzs[ROWS][nc] | dhs[ROWS][nc] | red[32]
__global__ void __launch_bounds__(THREADS) patch_mlp_bwd_rows_kernel(
    const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ h,
    const int64_t* __restrict__ ids, int N, int C, int64_t F, int P, const float* __restrict__ W1,
    const float* __restrict__ W2, int nc, float* __restrict__ dz, float* __restrict__ dh, float* __restrict__ dfeat) {
  gb_pdl_enter();
  extern __shared__ float sh[];
  float* dzs = sh;
  float* dhs = dzs + ROWS * nc;
  float* red = dhs + ROWS * nc;
  const int R = N * P;
  const int r0 = blockIdx.x * ROWS;
  const int tid = threadIdx.x;
  for (int rr = 0; rr < ROWS; ++rr) {
    const int r = r0 + rr;
    float s2 = 0.f, sd = 0.f;
    if (r < R)
      for (int j = tid; j < nc; j += THREADS) {
        const float zv = z[(int64_t)r * nc + j];
        s2 += zv * zv;
        sd += zv * dy[(int64_t)r * nc + j];
      }
    s2 = block_sum(s2, red);
    sd = block_sum(sd, red);
    const float nrm = sqrtf(s2);
    const float inv = 1.f / (nrm + 1e-7f);
    // y = z * inv, inv = 1 / (n + eps):  dz = dy * inv - z * (dy . z) * inv^2 / n
    const float coef = nrm > 0.f ? sd * inv * inv / nrm : 0.f;
    for (int j = tid; j < nc; j += THREADS) {
      float v = 0.f;
      if (r < R) {
        v = dy[(int64_t)r * nc + j] * inv - z[(int64_t)r * nc + j] * coef;
        dz[(int64_t)r * nc + j] = v;
      }
      dzs[rr * nc + j] = v;
    }
  }
  __syncthreads();
  // dh[r][k] = (h > 0) * sum_j dz[r][j] * W2[j][k]   (thread k: W2 read coalesced over k)
  for (int k = tid; k < nc; k += THREADS) {
    float acc[ROWS];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) acc[rr] = 0.f;
    for (int j = 0; j < nc; ++j) {
      const float wv = __ldg(W2 + (int64_t)j * nc + k);
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) acc[rr] = fmaf(dzs[rr * nc + j], wv, acc[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
      const int r = r0 + rr;
      float v = 0.f;
      if (r < R) {
        v = h[(int64_t)r * nc + k] > 0.f ? acc[rr] : 0.f;
        dh[(int64_t)r * nc + k] = v;
      }
      dhs[rr * nc + k] = v;
    }
  }
  __syncthreads();
  if (dfeat == nullptr) return;
  // dx[r][c] = sum_k dh[r][k] * W1[k][c]
  for (int c = tid; c < C; c += THREADS) {
    float acc[ROWS];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) acc[rr] = 0.f;
    for (int k = 0; k < nc; ++k) {
      const float wv = __ldg(W1 + (int64_t)k * C + c);
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) acc[rr] = fmaf(dhs[rr * nc + k], wv, acc[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
      const int r = r0 + rr;
      if (r < R) {
        const int n = r / P, p = r - n * P;
        dfeat[((int64_t)n * C + c) * F + ids[p]] = acc[rr];
      }
    }
  }
}

// Parameter gradients: dW[j][k] = sum_r g[r][j] * a[r][k] (g = dz, a = h for W2; g = dh, a = x for W1) and
// db[j] = sum_r g[r][j].  One CTA per JT output rows j; thread k owns columns k, k + 256, ...; rows r in a fixed
// order (deterministic, no atomics).
constexpr int JT = 4;
__global__ void __launch_bounds__(THREADS) patch_mlp_bwd_params_kernel(const float* __restrict__ g,
                                                                       const float* __restrict__ a, int R, int nj,
                                                                       int nk, float* __restrict__ dW,
                                                                       float* __restrict__ db) {
  gb_pdl_enter();
  const int j0 = blockIdx.x * JT;
  const int tid = threadIdx.x;
  for (int k = tid; k < nk; k += THREADS) {
    float acc[JT];
#pragma unroll
    for (int jj = 0; jj < JT; ++jj) acc[jj] = 0.f;
    for (int r = 0; r < R; ++r) {
      const float av = __ldg(a + (int64_t)r * nk + k);
#pragma unroll
      for (int jj = 0; jj < JT; ++jj)
        if (j0 + jj < nj) acc[jj] = fmaf(__ldg(g + (int64_t)r * nj + j0 + jj), av, acc[jj]);
    }
#pragma unroll
    for (int jj = 0; jj < JT; ++jj)
      if (j0 + jj < nj) dW[(int64_t)(j0 + jj) * nk + k] = acc[jj];
  }
  if (db != nullptr && tid < JT && j0 + tid < nj) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += g[(int64_t)r * nj + j0 + tid];
    db[j0 + tid] = s;
  }
}

}  // namespace

extern "C" int gb_patch_mlp_fwd(const float* feat, const int64_t* ids, int N, int C, int64_t F, int P, const float* W1,
                                const float* b1, const float* W2, const float* b2, int nc, float* xg, float* h,
                                float* z, float* y, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GB_CHECK(feat && ids && W1 && b1 && W2 && b2 && xg && h && z && y, "gb_patch_mlp_fwd: null pointer");
  GB_CHECK(N >= 1 && C >= 1 && P >= 1 && nc >= 1 && F >= P, "gb_patch_mlp_fwd: bad sizes N=%d C=%d P=%d nc=%d", N, C, P, nc);
  const size_t smem = ((size_t)ROWS * C + 2 * (size_t)ROWS * nc + 32) * sizeof(float);
  GB_CHECK(smem <= 200 * 1024, "gb_patch_mlp_fwd: C=%d / nc=%d need %zu bytes of shared memory", C, nc, smem);
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    GB_CUDA(cudaFuncSetAttribute(patch_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = 200 * 1024;
  }
  const int R = N * P;
  gb_klaunch(patch_mlp_fwd_kernel, dim3(gb_cdiv(R, ROWS)), dim3(THREADS), smem, st, feat, ids, N, C, F, P, W1, b1, W2, b2,
             nc, xg, h, z, y);
  GB_LAUNCH_CHECK();
  return 0;
}

extern "C" int gb_patch_mlp_bwd(const float* dy, const float* xg, const float* h, const float* z, const int64_t* ids,
                                int N, int C, int64_t F, int P, const float* W1, const float* W2, int nc, float* dz,
                                float* dh, float* dfeat, float* dW1, float* db1, float* dW2, float* db2, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GB_CHECK(dy && xg && h && z && ids && W1 && W2 && dz && dh, "gb_patch_mlp_bwd: null pointer");
  GB_CHECK(N >= 1 && C >= 1 && P >= 1 && nc >= 1, "gb_patch_mlp_bwd: bad sizes");
  const size_t smem = (2 * (size_t)ROWS * nc + 32) * sizeof(float);
  GB_CHECK(smem <= 200 * 1024, "gb_patch_mlp_bwd: nc=%d needs %zu bytes of shared memory", nc, smem);
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    GB_CUDA(cudaFuncSetAttribute(patch_mlp_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = 200 * 1024;
  }
  const int R = N * P;
  gb_klaunch(patch_mlp_bwd_rows_kernel, dim3(gb_cdiv(R, ROWS)), dim3(THREADS), smem, st, dy, z, h, ids, N, C, F, P, W1, W2,
             nc, dz, dh, dfeat);
  GB_LAUNCH_CHECK();
  if (dW2 != nullptr) {
    gb_klaunch(patch_mlp_bwd_params_kernel, dim3(gb_cdiv(nc, JT)), dim3(THREADS), 0, st, (const float*)dz, h, R, nc, nc,
               dW2, db2);
    GB_LAUNCH_CHECK();
  }
  if (dW1 != nullptr) {
    gb_klaunch(patch_mlp_bwd_params_kernel, dim3(gb_cdiv(nc, JT)), dim3(THREADS), 0, st, (const float*)dh, xg, R, nc, C,
               dW1, db1);
    GB_LAUNCH_CHECK();
  }
  return 0;
}
