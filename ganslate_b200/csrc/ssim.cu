// SSIM distance loss and its analytic gradient -- ganslate/nn/losses/utils/ssim.py:51-99 (used by CycleLoss when
// proportion_ssim > 0, ganslate/nn/losses/cyclegan_losses.py:77-101).
//
//   X, Y: [planes][H][W] fp32 (planes = N*C; a 5-D input is viewed as (N*C) x D x H x W by the reference, i.e. depth
//   slices act as channels of the depthwise filter -> planes = N*C*D), both mapped by v = in * in_scale + in_shift
//   ((x + 1) / 2 in CycleLoss).  With the separable 11-tap Gaussian w (sigma 1.5, "valid" filtering):
//     mu1 = w*X, mu2 = w*Y, s1 = w*X^2 - mu1^2, s2 = w*Y^2 - mu2^2, s12 = w*XY - mu1 mu2
//     S1 = (2 mu1 mu2 + C1) / (mu1^2 + mu2^2 + C1),  S2 = (2 s12 + C2) / (s1 + s2 + C2)
//     loss = mean over the (H-10) x (W-10) map of sqrt(relu(2 - S1 - S2))
//
// The reference runs 10 depthwise cuDNN convolutions + ~20 elementwise kernels over full-size temporaries.  Here one
// CTA stages a 42x42 (forward) / 52x52 (backward) tile of X and Y in shared memory and runs both separable passes
// of all five moment images there; nothing but the loss scalar / the gradient image goes back to HBM.
//
// Gradient wrt X (Y is the real image; the reference's call order is ssim(reconstructed, real)):
//   dL/dX_q = in_scale * sum_p w(q - p) [ Gmu_p + 2 X_q Gxx_p + Y_q Gxy_p ]      (p over the valid map)
//   g_p = dloss / Np * 1 / (2 sqrt(S_p)) for S_p > 0 else 0   (the reference would produce 0 * inf = NaN at S_p == 0)
//   Gmu = -g (dS1/dmu1 + dS2/dmu1),  Gxx = g A2 / B2^2,  Gxy = -2 g / B2
#include "gb_common.cuh"

namespace {

constexpr int WIN = 11;
constexpr int R = WIN - 1;   // 10
constexpr int TS = 32;       // tile edge (outputs in forward, input pixels in backward)
constexpr int FT = TS + R;   // 42: forward input tile / backward coefficient tile
constexpr int BT = FT + R;   // 52: backward input tile

struct SsimArgs {
  const float* x;
  const float* y;
  int planes, H, W;
  float in_scale, in_shift, c1, c2, inv_count;
  float win[WIN];
};

__device__ __forceinline__ float block_sum256(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  v = (threadIdx.x < 8) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;
}

// the five moments at one map position from the vertical pass over hs[5][rows][cols]
struct Moments {
  float mu1, mu2, exx, eyy, exy;
};

__global__ void __launch_bounds__(256) ssim_fwd_kernel(const __grid_constant__ SsimArgs a, float* __restrict__ loss) {
  gb_pdl_enter();
  __shared__ float sx[FT][FT + 1], sy[FT][FT + 1];
  __shared__ float hs[5][FT][TS + 1];
  __shared__ float red[8];
  const int tid = threadIdx.x;
  const int y0 = blockIdx.y * TS, x0 = blockIdx.x * TS;
  const float* xp = a.x + (int64_t)blockIdx.z * a.H * a.W;
  const float* yp = a.y + (int64_t)blockIdx.z * a.H * a.W;
  for (int i = tid; i < FT * FT; i += 256) {
    const int r = i / FT, c = i - r * FT;
    const int gy = y0 + r, gx = x0 + c;
    float vx = 0.f, vy = 0.f;
    if (gy < a.H && gx < a.W) {
      vx = xp[(int64_t)gy * a.W + gx] * a.in_scale + a.in_shift;
      vy = yp[(int64_t)gy * a.W + gx] * a.in_scale + a.in_shift;
    }
    sx[r][c] = vx;
    sy[r][c] = vy;
  }
  __syncthreads();
  for (int i = tid; i < FT * TS; i += 256) {  // horizontal pass
    const int r = i / TS, c = i - r * TS;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
      const float w = a.win[k], p = sx[r][c + k], q = sy[r][c + k];
      s0 += w * p;
      s1 += w * q;
      s2 += w * p * p;
      s3 += w * q * q;
      s4 += w * p * q;
    }
    hs[0][r][c] = s0; hs[1][r][c] = s1; hs[2][r][c] = s2; hs[3][r][c] = s3; hs[4][r][c] = s4;
  }
  __syncthreads();
  float acc = 0.f;
  for (int i = tid; i < TS * TS; i += 256) {  // vertical pass + map
    const int r = i / TS, c = i - r * TS;
    if (y0 + r >= a.H - R || x0 + c >= a.W - R) continue;
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
      const float w = a.win[k];
#pragma unroll
      for (int j = 0; j < 5; ++j) m[j] += w * hs[j][r + k][c];
    }
    const float mu1 = m[0], mu2 = m[1];
    const float s1 = m[2] - mu1 * mu1, s2 = m[3] - mu2 * mu2, s12 = m[4] - mu1 * mu2;
    const float S1 = (2.f * mu1 * mu2 + a.c1) / (mu1 * mu1 + mu2 * mu2 + a.c1);
    const float S2 = (2.f * s12 + a.c2) / (s1 + s2 + a.c2);
    acc += sqrtf(fmaxf(2.f - (S1 + S2), 0.f));
  }
  acc = block_sum256(acc, red);
  if (tid == 0) atomicAdd(loss, acc * a.inv_count);
}

// dynamic smem layout (floats): sx[BT][BT+1], sy[BT][BT+1], G[3][FT][FT+1], hs[5][BT][FT+1] (reused as hg[3][FT][TS+1])
constexpr int SXP = BT + 1, GP = FT + 1, HP = FT + 1, HGP = TS + 1;
constexpr int BWD_SMEM_FLOATS = 2 * BT * SXP + 3 * FT * GP + 5 * BT * HP;

__global__ void __launch_bounds__(256) ssim_bwd_kernel(const __grid_constant__ SsimArgs a, const float* __restrict__ dloss,
                                                       float* __restrict__ grad) {
  gb_pdl_enter();
  extern __shared__ float sm[];
  float* sx = sm;
  float* sy = sx + BT * SXP;
  float* G = sy + BT * SXP;
  float* hs = G + 3 * FT * GP;
  float* hg = hs;  // the horizontal moments are dead once G is built
  const int tid = threadIdx.x;
  const int qy0 = blockIdx.y * TS, qx0 = blockIdx.x * TS;   // first input pixel of this tile
  const int iy0 = qy0 - R, ix0 = qx0 - R;                   // first input pixel staged = first map position needed
  const float* xp = a.x + (int64_t)blockIdx.z * a.H * a.W;
  const float* yp = a.y + (int64_t)blockIdx.z * a.H * a.W;
  const float gscale = dloss[0] * a.inv_count;
  for (int i = tid; i < BT * BT; i += 256) {
    const int r = i / BT, c = i - r * BT;
    const int gy = iy0 + r, gx = ix0 + c;
    float vx = 0.f, vy = 0.f;
    if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
      vx = xp[(int64_t)gy * a.W + gx] * a.in_scale + a.in_shift;
      vy = yp[(int64_t)gy * a.W + gx] * a.in_scale + a.in_shift;
    }
    sx[r * SXP + c] = vx;
    sy[r * SXP + c] = vy;
  }
  __syncthreads();
  for (int i = tid; i < BT * FT; i += 256) {  // horizontal pass of the five moments: 52 rows x 42 map columns
    const int r = i / FT, c = i - r * FT;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
      const float w = a.win[k], p = sx[r * SXP + c + k], q = sy[r * SXP + c + k];
      s0 += w * p;
      s1 += w * q;
      s2 += w * p * p;
      s3 += w * q * q;
      s4 += w * p * q;
    }
    hs[(0 * BT + r) * HP + c] = s0;
    hs[(1 * BT + r) * HP + c] = s1;
    hs[(2 * BT + r) * HP + c] = s2;
    hs[(3 * BT + r) * HP + c] = s3;
    hs[(4 * BT + r) * HP + c] = s4;
  }
  __syncthreads();
  for (int i = tid; i < FT * FT; i += 256) {  // coefficient maps at the 42 x 42 map positions p = (iy0 + r, ix0 + c)
    const int r = i / FT, c = i - r * FT;
    const int py = iy0 + r, px = ix0 + c;
    float gmu = 0.f, gxx = 0.f, gxy = 0.f;
    if (py >= 0 && py < a.H - R && px >= 0 && px < a.W - R) {
      float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < WIN; ++k) {
        const float w = a.win[k];
#pragma unroll
        for (int j = 0; j < 5; ++j) m[j] += w * hs[(j * BT + r + k) * HP + c];
      }
      const float mu1 = m[0], mu2 = m[1];
      const float s1 = m[2] - mu1 * mu1, s2 = m[3] - mu2 * mu2, s12 = m[4] - mu1 * mu2;
      const float A1 = 2.f * mu1 * mu2 + a.c1, B1 = mu1 * mu1 + mu2 * mu2 + a.c1;
      const float A2 = 2.f * s12 + a.c2, B2 = s1 + s2 + a.c2;
      const float S = 2.f - (A1 / B1 + A2 / B2);
      if (S > 0.f) {
        const float g = gscale * 0.5f * rsqrtf(S);
        const float dS1 = (2.f * mu2 * B1 - 2.f * mu1 * A1) / (B1 * B1);
        const float dS2 = -2.f * mu2 / B2 + 2.f * mu1 * A2 / (B2 * B2);
        gmu = -g * (dS1 + dS2);
        gxx = g * A2 / (B2 * B2);
        gxy = -2.f * g / B2;
      }
    }
    G[(0 * FT + r) * GP + c] = gmu;
    G[(1 * FT + r) * GP + c] = gxx;
    G[(2 * FT + r) * GP + c] = gxy;
  }
  __syncthreads();
  // transpose of the valid filter: input pixel q receives w(q - p) from the map positions p = q - k, k = 0..10.
  // Local: q = (r, c) in the 32 x 32 tile <-> map local (r + 10 - ky, c + 10 - kx).
  for (int i = tid; i < 3 * FT * TS; i += 256) {  // horizontal: 3 maps x 42 rows x 32 columns
    const int mth = i / (FT * TS), rem = i - mth * (FT * TS);
    const int r = rem / TS, c = rem - r * TS;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < WIN; ++k) s += a.win[k] * G[(mth * FT + r) * GP + c + R - k];
    hg[(mth * FT + r) * HGP + c] = s;
  }
  __syncthreads();
  float* gp = grad + (int64_t)blockIdx.z * a.H * a.W;
  for (int i = tid; i < TS * TS; i += 256) {
    const int r = i / TS, c = i - r * TS;
    const int qy = qy0 + r, qx = qx0 + c;
    if (qy >= a.H || qx >= a.W) continue;
    float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < WIN; ++k) {
      const float w = a.win[k];
#pragma unroll
      for (int mth = 0; mth < 3; ++mth) o[mth] += w * hg[(mth * FT + r + R - k) * HGP + c];
    }
    const float xq = sx[(r + R) * SXP + c + R], yq = sy[(r + R) * SXP + c + R];
    gp[(int64_t)qy * a.W + qx] = a.in_scale * (o[0] + 2.f * xq * o[1] + yq * o[2]);
  }
}

int fill_args(SsimArgs& a, const float* x, const float* y, int planes, int H, int W, float in_scale, float in_shift,
              float data_range) {
  GB_CHECK(x && y && planes > 0, "gb_ssim: bad arguments");
  GB_CHECK(H > R && W > R, "gb_ssim: image (%d x %d) smaller than the 11-tap window", H, W);
  GB_CHECK(planes <= 65535 * 32, "gb_ssim: too many planes");
  a.x = x;
  a.y = y;
  a.planes = planes;
  a.H = H;
  a.W = W;
  a.in_scale = in_scale;
  a.in_shift = in_shift;
  a.c1 = (0.01f * data_range) * (0.01f * data_range);   // K = (0.01, 0.03), ssim.py:53,80-81
  a.c2 = (0.03f * data_range) * (0.03f * data_range);
  a.inv_count = 1.f / ((float)planes * (float)(H - R) * (float)(W - R));
  // _fspecial_gauss_1d (ssim.py:22-41): size 11, sigma 1.5, fp32
  float g[WIN], sum = 0.f;
  for (int i = 0; i < WIN; ++i) {
    const float c = (float)(i - WIN / 2);
    g[i] = expf(-(c * c) / (2.f * 1.5f * 1.5f));
    sum += g[i];
  }
  for (int i = 0; i < WIN; ++i) a.win[i] = g[i] / sum;
  return 0;
}

}  // namespace

extern "C" int gb_ssim_fwd(const float* x, const float* y, int planes, int H, int W, float in_scale, float in_shift,
                           float data_range, float* loss, void* stream) {
  SsimArgs a;
  if (int r = fill_args(a, x, y, planes, H, W, in_scale, in_shift, data_range)) return r;
  GB_CHECK(loss != nullptr, "gb_ssim_fwd: null loss");
  dim3 grid(gb_cdiv(W - R, TS), gb_cdiv(H - R, TS), planes);
  GB_CHECK(grid.z <= 65535, "gb_ssim_fwd: too many planes (%d)", planes);
  gb_klaunch(ssim_fwd_kernel, grid, 256, 0, (cudaStream_t)stream, a, loss);
  GB_LAUNCH_CHECK();
  return 0;
}

extern "C" int gb_ssim_bwd(const float* x, const float* y, int planes, int H, int W, float in_scale, float in_shift,
                           float data_range, const float* dloss, float* grad_x, void* stream) {
  SsimArgs a;
  if (int r = fill_args(a, x, y, planes, H, W, in_scale, in_shift, data_range)) return r;
  GB_CHECK(dloss && grad_x, "gb_ssim_bwd: null pointer");
  static bool attr_set = false;
  const int smem = BWD_SMEM_FLOATS * (int)sizeof(float);
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(ssim_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(gb_cdiv(W, TS), gb_cdiv(H, TS), planes);
  GB_CHECK(grid.z <= 65535, "gb_ssim_bwd: too many planes (%d)", planes);
  gb_klaunch(ssim_bwd_kernel, grid, 256, smem, (cudaStream_t)stream, a, dloss, grad_x);
  GB_LAUNCH_CHECK();
  return 0;
}
