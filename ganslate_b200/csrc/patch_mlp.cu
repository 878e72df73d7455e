// CUT's FeaturePatchMLP as fused kernels (ganslate/nn/gans/unpaired/cut.py:229-294): gather `P` positions of a
// feature map (the same ids for every image), Linear(C -> nc) + ReLU + Linear(nc -> nc), L2-normalise each row.
//   x[r][c]  = feat[n][c][id[p]]                      r = n * P + p
//   h        = relu(x W1^T + b1)                      W1: (nc, C)  torch.nn.Linear layout
//   z        = h W2^T + b2                            W2: (nc, nc)
//   y        = z / (||z||_2 + 1e-7)
// The problem is tiny (P = 256 rows per image, C <= 256, nc = 256: 33 MFLOP per feature and batch item) and sits
// between two tensor-core passes; fp32 CUDA-core FMAs keep it bit-comparable with the fp32 reference (no TF32, no
// library GEMM) and the whole forward is ONE launch per feature, the backward three.  Every CTA streams the weights
// once from L2 straight into registers (16-byte loads along the contiguous dimension of W, whichever side of the
// product that is -- see gemm_reduce_lanes / gemm_reduce_warps); the eight rows of activations stay on chip.
#include "gb_common.cuh"

namespace {

// Four rows of x per CTA: at batch 1 (P = 256 rows) that is 64 CTAs, each streaming W1 and W2 once from L2 -- with eight
// rows the 32 CTAs were bound by their own instruction issue (64 FMAs + the cross-lane sum per row of W and warp), with
// two the L2 -> SM traffic (every CTA reads all of W) would bound instead.
constexpr int ROWS = 4;
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int KB = 256;       // columns of W one warp covers at a time: 8 consecutive per lane
constexpr int GATHER_U = 4;
static_assert(ROWS == 4 && ROWS <= WARPS, "warp_sum4 and the float4 row vectors assume four rows");

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Eight consecutive floats of a row-major matrix row, zero beyond `n_valid` columns.  VEC: the row starts are 16-byte
// aligned (row length % 4 == 0), hence n_valid % 4 == 0 and each half is one 16-byte load or nothing.
template <bool VEC>
__device__ __forceinline__ void load8(const float* __restrict__ p, int n_valid, float (&w)[8]) {
  if constexpr (VEC) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (n_valid >= 4) a = __ldg(reinterpret_cast<const float4*>(p));
    if (n_valid >= 8) b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w, w[4] = b.x, w[5] = b.y, w[6] = b.z, w[7] = b.w;
  } else {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      w[u] = 0.f;
      if (u < n_valid) w[u] = __ldg(p + u);
    }
  }
}

// Sum v[0..3] over the 32 lanes; afterwards every lane holds in v[0] the total of entry 2 * bit4(lane) + bit3(lane)
// (halving exchange: 2 + 1 + 1 + 1 + 1 shuffles instead of 4 x 5).
__device__ __forceinline__ void warp_sum4(float (&v)[ROWS], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b4 ? v[i] : v[i + 2], keep = b4 ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  {
    const float send = b3 ? v[0] : v[1], keep = b3 ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 4);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// out[rr][j] = sum_k in[rr][k] * W[j][k] for the CTA's rows; W row-major (nj, K) as torch.nn.Linear stores it.
// The reduction dimension is the contiguous one, so the LANES split k (8 consecutive columns each: a warp reads a whole
// 1 KB row segment of W with two 16-byte loads per lane, straight from L2 into registers, no staging and no barrier),
// the warps split j, the inputs of a lane's columns live in registers, and a row of products is summed across the lanes
// with warp_sum4.  Rows of W travel in batches of JU per warp, double-buffered in registers: the next batch is in flight
// while the current one is used (all warps of all CTAs start in lockstep: without this the SM alternates between
// waiting for L2 and computing).  `in` / `out` are shared memory, [ROWS][K] and [ROWS][nj]; K > 256 adds into `out`
// block by block (each element has one owner lane).  No bias; the caller follows up.
constexpr int JU = 4;
template <bool VEC, bool ACCUM>
__device__ __forceinline__ void gemm_reduce_lanes_block(const float* __restrict__ W, int K, int nj, int kb, const float* in,
                                                        float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int orow = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
  const int k0 = kb + lane * 8;
  const int n_valid = max(0, min(8, K - k0));
  float xr[ROWS][8];
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      xr[rr][u] = 0.f;
      if (u < n_valid) xr[rr][u] = in[rr * K + k0 + u];
    }
  const float* wl = W + k0;
  constexpr int STEP = WARPS * JU;
  auto fetch = [&](float (&w)[JU][8], int j0) {
#pragma unroll
    for (int q = 0; q < JU; ++q)   // (clamped: beyond the last row the last row is read again and its result dropped)
      load8<VEC>(wl + (int64_t)min(j0 + q * WARPS, nj - 1) * K, n_valid, w[q]);
  };
  auto use = [&](const float (&w)[JU][8], int j0) {
#pragma unroll
    for (int q = 0; q < JU; ++q) {
      const int j = j0 + q * WARPS;
      float v[ROWS];
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) {
        float a = xr[rr][0] * w[q][0];
#pragma unroll
        for (int u = 1; u < 8; ++u) a = fmaf(xr[rr][u], w[q][u], a);
        v[rr] = a;
      }
      warp_sum4(v, lane);
      if ((lane & 7) == 0 && j < nj) {
        if constexpr (ACCUM) out[orow * nj + j] += v[0];
        else out[orow * nj + j] = v[0];
      }
    }
  };
  float wa[JU][8], wb[JU][8];
  fetch(wa, warp);
  for (int j0 = warp; j0 < nj; j0 += 2 * STEP) {
    fetch(wb, j0 + STEP);
    use(wa, j0);
    fetch(wa, j0 + 2 * STEP);
    use(wb, j0 + STEP);
  }
}

__device__ __forceinline__ void gemm_reduce_lanes(const float* __restrict__ W, int K, int nj, const float* in, float* out) {
  if ((K & 3) == 0) {
    gemm_reduce_lanes_block<true, false>(W, K, nj, 0, in, out);
    for (int kb = KB; kb < K; kb += KB) gemm_reduce_lanes_block<true, true>(W, K, nj, kb, in, out);
  } else {
    gemm_reduce_lanes_block<false, false>(W, K, nj, 0, in, out);
    for (int kb = KB; kb < K; kb += KB) gemm_reduce_lanes_block<false, true>(W, K, nj, kb, in, out);
  }
}

// out[rr][c] = sum_r inT[r][rr] * W[r][c]; W row-major (nr, ncol).  Here the OUTPUT dimension is the contiguous one:
// the lanes split c (8 consecutive columns each, ROWS x 8 accumulators per lane), the warps split the reduction (warp w
// takes r = w, w + 8, ...; inT[r][0..3] is one broadcast 16-byte shared load per 32 FMAs; batches of RU rows of W
// double-buffered in registers as above), and the eight partial tiles are summed through shared memory in warp order
// (deterministic).  `part`: [WARPS][ROWS][KB] floats.  One 256-column block `cb` per call; afterwards thread t holds
// out[0..3] of column cb + t.
constexpr int RU = 4;
template <bool VEC>
__device__ __forceinline__ void gemm_reduce_warps_t(const float* __restrict__ W, int nr, int ncol, int cb, const float* inT,
                                                    float* part, float (&out)[ROWS]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = cb + lane * 8;
  const int n_valid = max(0, min(8, ncol - c0));
  const float4* in4 = reinterpret_cast<const float4*>(inT);
  const float* wl = W + c0;
  float acc[ROWS][8];
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[rr][u] = 0.f;
  constexpr int STEP = WARPS * RU;
  auto fetch = [&](float (&w)[RU][8], int rb) {
#pragma unroll
    for (int q = 0; q < RU; ++q) {
      const int r = rb + q * WARPS;
      load8<VEC>(wl + (int64_t)min(r, nr - 1) * ncol, r < nr ? n_valid : 0, w[q]);   // beyond the last row: zeros
    }
  };
  auto use = [&](const float (&w)[RU][8], int rb) {
#pragma unroll
    for (int q = 0; q < RU; ++q) {
      const float4 a = in4[min(rb + q * WARPS, nr - 1)];
      const float x[ROWS] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[rr][u] = fmaf(x[rr], w[q][u], acc[rr][u]);
    }
  };
  float wa[RU][8], wb[RU][8];
  fetch(wa, warp);
  for (int rb = warp; rb < nr; rb += 2 * STEP) {
    fetch(wb, rb + STEP);
    use(wa, rb);
    fetch(wa, rb + 2 * STEP);
    use(wb, rb + STEP);
  }
  __syncthreads();   // `part` is free again (previous block's sums have been read)
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    float4* dst = reinterpret_cast<float4*>(part + ((size_t)warp * ROWS + rr) * KB + lane * 8);
    dst[0] = make_float4(acc[rr][0], acc[rr][1], acc[rr][2], acc[rr][3]);
    dst[1] = make_float4(acc[rr][4], acc[rr][5], acc[rr][6], acc[rr][7]);
  }
  __syncthreads();
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += part[((size_t)w * ROWS + rr) * KB + threadIdx.x];
    out[rr] = s;
  }
}

__device__ __forceinline__ void gemm_reduce_warps(const float* __restrict__ W, int nr, int ncol, int cb, const float* inT,
                                                  float* part, float (&out)[ROWS]) {
  if ((ncol & 3) == 0) gemm_reduce_warps_t<true>(W, nr, ncol, cb, inT, part, out);
  else gemm_reduce_warps_t<false>(W, nr, ncol, cb, inT, part, out);
}

// grid = ceil(R / ROWS); smem: xs[ROWS][C] | hs[ROWS][nc] | zs[ROWS][nc]
__global__ void __launch_bounds__(THREADS) patch_mlp_fwd_kernel(
    const float* __restrict__ feat, const int64_t* __restrict__ ids, int N, int C, int64_t F, int P,
    const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
    const float* __restrict__ b2, int nc, float* __restrict__ xg, float* __restrict__ h, float* __restrict__ z,
    float* __restrict__ y) {
  gb_pdl_enter();
  extern __shared__ __align__(16) float sh[];
  float* xs = sh;
  float* hs = xs + (size_t)ROWS * C;
  float* zs = hs + (size_t)ROWS * nc;
  const int R = N * P;
  const int r0 = blockIdx.x * ROWS;
  const int tid = threadIdx.x;
  // gather: index i = rr * C + c; all loads of a batch are issued before the first use (two dependent round trips per
  // element otherwise, one after the other)
  for (int base = tid; base < ROWS * C; base += THREADS * GATHER_U) {
    float v[GATHER_U];
    int64_t src[GATHER_U];
#pragma unroll
    for (int u = 0; u < GATHER_U; ++u) {
      const int i = base + u * THREADS;
      const int rr = i / C, c = i - rr * C, r = r0 + rr;
      src[u] = -1;
      if (i < ROWS * C && r < R) {
        const int n = r / P, p = r - n * P;
        src[u] = ((int64_t)n * C + c) * F + __ldg(ids + p);
      }
    }
#pragma unroll
    for (int u = 0; u < GATHER_U; ++u) v[u] = src[u] >= 0 ? __ldg(feat + src[u]) : 0.f;
#pragma unroll
    for (int u = 0; u < GATHER_U; ++u) {
      const int i = base + u * THREADS;
      if (i < ROWS * C) {
        xs[i] = v[u];
        if (src[u] >= 0) xg[(int64_t)r0 * C + i] = v[u];
      }
    }
  }
  __syncthreads();
  gemm_reduce_lanes(W1, C, nc, xs, hs);
  __syncthreads();
  for (int i = tid; i < ROWS * nc; i += THREADS) {
    const int rr = i / nc, j = i - rr * nc;
    const float v = fmaxf(hs[i] + b1[j], 0.f);
    hs[i] = v;
    if (r0 + rr < R) h[(int64_t)r0 * nc + i] = v;
  }
  __syncthreads();
  gemm_reduce_lanes(W2, nc, nc, hs, zs);
  __syncthreads();
  if (tid < ROWS * 32) {  // bias + L2 normalisation: warp rr owns row rr
    const int rr = tid >> 5, lane = tid & 31;
    const int r = r0 + rr;
    float s = 0.f;
    for (int j = lane; j < nc; j += 32) {
      const float v = zs[rr * nc + j] + b2[j];
      zs[rr * nc + j] = v;
      s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    const float inv = 1.f / (sqrtf(s) + 1e-7f);
    if (r < R)
      for (int j = lane; j < nc; j += 32) {
        const float v = zs[rr * nc + j];
        z[(int64_t)r * nc + j] = v;
        y[(int64_t)r * nc + j] = v * inv;
      }
  }
}

// Row-wise part of the backward: dz (L2-norm), dh (ReLU mask), dx scattered into dfeat (zero-initialised by the
// caller; ids are distinct within a feature, so no two rows of one image meet).  grid = ceil(R / ROWS).
// smem: dzsT[nc][ROWS] | dhsT[nc][ROWS] | part[WARPS][ROWS][KB]
__global__ void __launch_bounds__(THREADS) patch_mlp_bwd_rows_kernel(
    const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ h,
    const int64_t* __restrict__ ids, int N, int C, int64_t F, int P, const float* __restrict__ W1,
    const float* __restrict__ W2, int nc, float* __restrict__ dz, float* __restrict__ dh, float* __restrict__ dfeat) {
  gb_pdl_enter();
  extern __shared__ __align__(16) float sh[];
  float* dzsT = sh;
  float* dhsT = dzsT + (size_t)nc * ROWS;
  float* part = dhsT + (size_t)nc * ROWS;
  const int R = N * P;
  const int r0 = blockIdx.x * ROWS;
  const int tid = threadIdx.x;
  if (tid < ROWS * 32) {  // warp rr owns row rr
    const int rr = tid >> 5, lane = tid & 31;
    const int r = r0 + rr;
    float s2 = 0.f, sd = 0.f;
    if (r < R)
      for (int j = lane; j < nc; j += 32) {
        const float zv = z[(int64_t)r * nc + j];
        s2 = fmaf(zv, zv, s2);
        sd = fmaf(zv, dy[(int64_t)r * nc + j], sd);
      }
    s2 = warp_sum(s2);
    sd = warp_sum(sd);
    const float nrm = sqrtf(s2);
    const float inv = 1.f / (nrm + 1e-7f);
    // y = z * inv, inv = 1 / (n + eps):  dz = dy * inv - z * (dy . z) * inv^2 / n
    const float coef = nrm > 0.f ? sd * inv * inv / nrm : 0.f;
    for (int j = lane; j < nc; j += 32) {
      float v = 0.f;
      if (r < R) {
        v = dy[(int64_t)r * nc + j] * inv - z[(int64_t)r * nc + j] * coef;
        dz[(int64_t)r * nc + j] = v;
      }
      dzsT[j * ROWS + rr] = v;
    }
  }
  __syncthreads();
  // dh[r][k] = (h > 0) * sum_j dz[r][j] * W2[j][k]
  for (int cb = 0; cb < nc; cb += KB) {
    float out[ROWS];
    gemm_reduce_warps(W2, nc, nc, cb, dzsT, part, out);
    const int k = cb + tid;
    if (k < nc) {
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) {
        const int r = r0 + rr;
        float v = 0.f;
        if (r < R) {
          v = h[(int64_t)r * nc + k] > 0.f ? out[rr] : 0.f;
          dh[(int64_t)r * nc + k] = v;
        }
        dhsT[k * ROWS + rr] = v;
      }
    }
  }
  if (dfeat == nullptr) return;
  __syncthreads();
  // dx[r][c] = sum_k dh[r][k] * W1[k][c]
  for (int cb = 0; cb < C; cb += KB) {
    float out[ROWS];
    gemm_reduce_warps(W1, nc, C, cb, dhsT, part, out);
    const int c = cb + tid;
    if (c < C) {
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr) {
        const int r = r0 + rr;
        if (r < R) {
          const int n = r / P, p = r - n * P;
          dfeat[((int64_t)n * C + c) * F + ids[p]] = out[rr];
        }
      }
    }
  }
}

// Parameter gradients: dW[j][k] = sum_r g[r][j] * a[r][k] (g = dz, a = h for W2; g = dh, a = x for W1) and
// db[j] = sum_r g[r][j].  A CTA owns a TJ x TK tile of dW and walks the rows r in chunks of RC staged in shared
// memory (next chunk prefetched into registers: batch 1 is two chunks, i.e. two L2 round trips); thread (kk, jg) holds
// 4 rows j of column kk.  Rows r in a fixed order: deterministic, no atomics.
// grid = (ceil(nj / TJ), ceil(nk / TK)).
constexpr int TJ = 16, TK = 64, RC = 128;
__global__ void __launch_bounds__(THREADS) patch_mlp_bwd_params_kernel(const float* __restrict__ g,
                                                                       const float* __restrict__ a, int R, int nj,
                                                                       int nk, float* __restrict__ dW,
                                                                       float* __restrict__ db) {
  gb_pdl_enter();
  __shared__ __align__(16) float gs[RC][TJ];
  __shared__ float as[RC][TK];
  const int tid = threadIdx.x;
  const int j0 = blockIdx.x * TJ, k0 = blockIdx.y * TK;
  const int kk = tid & (TK - 1), jg = tid >> 6;
  constexpr int NG = RC * TJ / THREADS, NA = RC * TK / THREADS;
  float pg[NG], pa[NA];
  auto fetch = [&](int rbase) {
#pragma unroll
    for (int t = 0; t < NG; ++t) {
      const int idx = tid + t * THREADS, rr = idx / TJ, jj = idx - rr * TJ;
      pg[t] = 0.f;
      if (rbase + rr < R && j0 + jj < nj) pg[t] = __ldg(g + (int64_t)(rbase + rr) * nj + j0 + jj);
    }
#pragma unroll
    for (int t = 0; t < NA; ++t) {
      const int idx = tid + t * THREADS, rr = idx / TK, k2 = idx - rr * TK;
      pa[t] = 0.f;
      if (rbase + rr < R && k0 + k2 < nk) pa[t] = __ldg(a + (int64_t)(rbase + rr) * nk + k0 + k2);
    }
  };
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, bacc[4] = {0.f, 0.f, 0.f, 0.f};
  fetch(0);
  for (int rbase = 0; rbase < R; rbase += RC) {
    __syncthreads();
#pragma unroll
    for (int t = 0; t < NG; ++t) (&gs[0][0])[tid + t * THREADS] = pg[t];
#pragma unroll
    for (int t = 0; t < NA; ++t) (&as[0][0])[tid + t * THREADS] = pa[t];
    __syncthreads();
    if (rbase + RC < R) fetch(rbase + RC);
#pragma unroll 16
    for (int rr = 0; rr < RC; ++rr) {
      const float av = as[rr][kk];
      const float4 g4 = *reinterpret_cast<const float4*>(&gs[rr][jg * 4]);
      acc[0] = fmaf(g4.x, av, acc[0]);
      acc[1] = fmaf(g4.y, av, acc[1]);
      acc[2] = fmaf(g4.z, av, acc[2]);
      acc[3] = fmaf(g4.w, av, acc[3]);
      bacc[0] += g4.x;
      bacc[1] += g4.y;
      bacc[2] += g4.z;
      bacc[3] += g4.w;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = j0 + jg * 4 + i;
    if (j < nj) {
      if (k0 + kk < nk) dW[(int64_t)j * nk + k0 + kk] = acc[i];
      if (db != nullptr && blockIdx.y == 0 && kk == 0) db[j] = bacc[i];
    }
  }
}

}  // namespace

extern "C" int gb_patch_mlp_fwd(const float* feat, const int64_t* ids, int N, int C, int64_t F, int P, const float* W1,
                                const float* b1, const float* W2, const float* b2, int nc, float* xg, float* h,
                                float* z, float* y, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GB_CHECK(feat && ids && W1 && b1 && W2 && b2 && xg && h && z && y, "gb_patch_mlp_fwd: null pointer");
  GB_CHECK(N >= 1 && C >= 1 && P >= 1 && nc >= 1 && F >= P, "gb_patch_mlp_fwd: bad sizes N=%d C=%d P=%d nc=%d", N, C, P, nc);
  const size_t smem = ((size_t)ROWS * C + 2 * (size_t)ROWS * nc) * sizeof(float);
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
  const size_t smem = (2 * (size_t)ROWS * nc + (size_t)WARPS * ROWS * KB) * sizeof(float);
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
    gb_klaunch(patch_mlp_bwd_params_kernel, dim3(gb_cdiv(nc, TJ), gb_cdiv(nc, TK)), dim3(THREADS), 0, st, (const float*)dz, h, R, nc, nc,
               dW2, db2);
    GB_LAUNCH_CHECK();
  }
  if (dW1 != nullptr) {
    gb_klaunch(patch_mlp_bwd_params_kernel, dim3(gb_cdiv(nc, TJ), gb_cdiv(C, TK)), dim3(THREADS), 0, st, (const float*)dh, xg, R, nc, C,
               dW1, db1);
    GB_LAUNCH_CHECK();
  }
  return 0;
}
