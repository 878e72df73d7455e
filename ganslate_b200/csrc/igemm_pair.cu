// Pair kernel: the implicit-GEMM convolution for the layers that dominate a step (stride-1 k x k convolutions with
// >= 64 input and output channels: the 18 residual-block convolutions of Resnet2D and their data gradients, the
// stride-1 4x4 layers of PatchGAN).
//
// Why: igemm_tma.cu moves A[128 px x 64 ch] (16 KB) + B[BN x 64] (32 KB at BN = 256) from L2 for every 512 cycles
// of tcgen05.mma -- 96 B/clk/SM against the ~42 B/clk/SM the L2 can deliver to 148 SMs (6300 B/clk, measured;
// the 256-CTA residual-block launch ran at exactly that rate: 256 MB per wave in 22 us).  Here a CTA owns TWO
// 16x8-pixel patches and two TMEM accumulators, so every weight tile B is loaded once for 2 x 128 rows, and the
// activations of a patch are loaded once per 64-channel chunk as a halo box {64 ch, kw+7 px, kh+15 rows}; the A
// operand of tap (ry, rx) is the same smem image read through a descriptor shifted by (ry*pitch + rx) pixels
// (igemm_halo.cu established that the 128-byte swizzle is a function of the absolute smem address, so a
// 128-byte-aligned start is enough).  Per chunk of a 3x3 convolution: 2 x 23 KB + 9 x 32 KB for 9216 MMA cycles
// = 36 B/clk/SM -- under the L2 ceiling, so the tensor pipe becomes the bound.
//
// Warp roles: warp 0 lane 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), all 8 warps = epilogue.
#include <cuda.h>
#include <string.h>
#include "gb_common.cuh"
#include "gb_geometry.h"
#include "gb_epilogue.cuh"
#include "gb_tma.h"

namespace {

constexpr int BK = 64;
constexpr int TW = 8, TH = 16;   // one patch = 16 rows x 8 columns of output pixels = 128 GEMM rows
constexpr int NSUB = 2;          // patches (accumulators) per CTA
constexpr int MAX_B_STAGES = 8;
constexpr int MAX_A_STAGES = 2;
constexpr int SMEM_LIMIT = 227 * 1024 - 2048;  // dynamic part: leaves room for the static bias_s

struct PairGeom {
  gb_fastdiv tiles_x, tiles_y, tiles_z;  // patch index -> (n, z, ty, tx)
  int nsub;                              // patches in the launch
  int hw, hh;                            // halo pitch (pixels per halo row) and rows
  int a_sub_bytes;                       // hw * hh * 128: bytes one halo box delivers
  int a_sub_stride;                      // a_sub_bytes rounded up to 1 KB (swizzle atoms stay aligned)
  int a_stages, b_stages;
  int ngroups;                           // distinct dz values (1 for 2-D)
  int dy_min, dx_min;
  int8_t group_dz[16];
  int16_t group_begin[17];               // taps are sorted by dz: taps [group_begin[g], group_begin[g+1]) share dz
};

__device__ __forceinline__ uint64_t make_smem_desc_kmajor(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                              // LBO (unused for swizzled K-major)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // stride between 8-row groups = one halo row
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B (base offset 0: see igemm_halo.cu)
  return d;
}

template <int BN>
__global__ void __launch_bounds__(256, 1)
igemm_pair_kernel(const __grid_constant__ gb_conv_params p, const __grid_constant__ CUtensorMap map_a,
                  const __grid_constant__ CUtensorMap map_b, const __grid_constant__ PairGeom pg) {
  gb_pdl_enter();
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int TMEM_COLS = NSUB * BN;  // 128 / 256 / 512
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int a_stage_bytes = NSUB * pg.a_sub_stride;
  const uint32_t a_base = base;
  const uint32_t b_base = base + pg.a_stages * a_stage_bytes;
  uint8_t* tail = smem + pg.a_stages * a_stage_bytes + pg.b_stages * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  // a_full[2], a_empty[2], b_full[8], b_empty[8], accum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 256);
  int8_t* taps_s = reinterpret_cast<int8_t*>(tail + 320);
  __shared__ float bias_s[BN];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const gb_conv_class& cc = p.cls[0];
  int q[3];
  gb_class_extents(p, 0, q);
  // the two patches of this CTA
  int x0[NSUB], y0[NSUB], z0[NSUB], nn[NSUB];
  bool valid[NSUB];
#pragma unroll
  for (int j = 0; j < NSUB; ++j) {
    uint32_t t = blockIdx.x * NSUB + j;
    valid[j] = t < (uint32_t)pg.nsub;
    uint32_t u = gb_div(t, pg.tiles_x);
    x0[j] = (int)(t - u * pg.tiles_x.d) * TW;
    t = u;
    u = gb_div(t, pg.tiles_y);
    y0[j] = (int)(t - u * pg.tiles_y.d) * TH;
    t = u;
    u = gb_div(t, pg.tiles_z);
    z0[j] = (int)(t - u * pg.tiles_z.d);
    nn[j] = (int)u;
  }
  const int nvalid = (valid[0] ? 1 : 0) + (valid[1] ? 1 : 0);
  const int n0 = blockIdx.y * BN;
  const int chunks = p.in.C >> 6;

  const uint32_t a_full = smem_u32(bars), a_empty = smem_u32(bars + MAX_A_STAGES);
  const uint32_t b_full = smem_u32(bars + 2 * MAX_A_STAGES), b_empty = smem_u32(bars + 2 * MAX_A_STAGES + MAX_B_STAGES);
  const uint32_t accum_bar = smem_u32(bars + 2 * MAX_A_STAGES + 2 * MAX_B_STAGES);
  if (tid == 0) {
    for (int s = 0; s < pg.a_stages; ++s) {
      mbar_init(a_full + 8 * s, 1);
      mbar_init(a_empty + 8 * s, 1);
    }
    for (int s = 0; s < pg.b_stages; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
  for (int i = tid; i < cc.ntaps; i += 256)
    *reinterpret_cast<uint32_t*>(taps_s + 4 * i) = *reinterpret_cast<const uint32_t*>(p.taps[cc.tap_begin + i]);
  for (int i = tid; i < BN; i += 256) bias_s[i] = (p.bias != nullptr && n0 + i < p.ncols) ? p.bias[n0 + i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one lane, see gb_elect_one)
    if (gb_elect_one()) {
      int as = 0, ait = 0, bs = 0, bit = 0;
      for (int g = 0; g < pg.ngroups; ++g) {
        const int dz = pg.group_dz[g];
        for (int c = 0; c < chunks; ++c) {
          if (ait > 0) mbar_wait(a_empty + 8 * as, (ait - 1) & 1);
          mbar_expect_tx(a_full + 8 * as, (uint32_t)(nvalid * pg.a_sub_bytes));
#pragma unroll
          for (int j = 0; j < NSUB; ++j)
            if (valid[j])
              tma_load_5d(a_base + as * a_stage_bytes + j * pg.a_sub_stride, &map_a, a_full + 8 * as, c * 64,
                          x0[j] + pg.dx_min, y0[j] + pg.dy_min, z0[j] + dz, nn[j]);
          for (int tl = pg.group_begin[g]; tl < pg.group_begin[g + 1]; ++tl) {
            if (bit > 0) mbar_wait(b_empty + 8 * bs, (bit - 1) & 1);
            mbar_expect_tx(b_full + 8 * bs, (uint32_t)B_BYTES);
            tma_load_2d(b_base + bs * B_BYTES, &map_b, b_full + 8 * bs, tl * p.in.C + c * 64, n0);
            if (++bs == pg.b_stages) {
              bs = 0;
              ++bit;
            }
          }
          if (++as == pg.a_stages) {
            as = 0;
            ++ait;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BN, 0, 0);
    int as = 0, ait = 0, bs = 0, bit = 0;
    uint32_t fresh = 1;  // the first MMA of each accumulator overwrites
    for (int g = 0; g < pg.ngroups; ++g) {
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(a_full + 8 * as, ait & 1);
        for (int tl = pg.group_begin[g]; tl < pg.group_begin[g + 1]; ++tl) {
          mbar_wait(b_full + 8 * bs, bit & 1);
          tc_fence_after();
          if (gb_elect_one()) {
            const int ry = taps_s[4 * tl + 1] - pg.dy_min, rx = taps_s[4 * tl + 2] - pg.dx_min;
            const uint64_t bdesc = make_smem_desc(b_base + bs * B_BYTES, 16, 1024);
#pragma unroll
            for (int j = 0; j < NSUB; ++j) {
              if (valid[j]) {
                const uint32_t a_s =
                    a_base + as * a_stage_bytes + j * pg.a_sub_stride + (uint32_t)(ry * pg.hw + rx) * 128u;
                const uint64_t adesc = make_smem_desc_kmajor(a_s, (uint32_t)pg.hw * 128u);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  umma_bf16(tmem_base + (uint32_t)(j * BN), adesc + 2 * k, bdesc + 2 * k, idesc, (fresh && k == 0) ? 0u : 1u);
              }
            }
            fresh = 0;
            umma_commit(b_empty + 8 * bs);
          }
          __syncwarp();
          if (++bs == pg.b_stages) {
            bs = 0;
            ++bit;
          }
        }
        if (lane == 0) umma_commit(a_empty + 8 * as);
        __syncwarp();
        if (++as == pg.a_stages) {
          as = 0;
          ++ait;
        }
      }
    }
    if (lane == 0) umma_commit(accum_bar);
    __syncwarp();
  }

  // -------------------------------------------------------------------- epilogue (all warps, one patch at a time)
  mbar_wait(accum_bar, 0);
  tc_fence_after();
#pragma unroll
  for (int j = 0; j < NSUB; ++j) {
    if (!valid[j]) continue;
    const int row = (warp & 3) * 32 + lane;
    const int h = row >> 3, w = row & 7;
    const int qy = y0[j] + h, qx = x0[j] + w;
    const bool row_ok = qy < q[1] && qx < q[2];
    int64_t ooff = 0;
    if (row_ok)
      ooff = gb_pix_offset(p.out, nn[j], z0[j] * p.out_mul[0] + cc.off[0], qy * p.out_mul[1] + cc.off[1],
                           qx * p.out_mul[2] + cc.off[2]);
    gb_conv_epilogue<BN>(p, tmem_base + (uint32_t)(j * BN), warp, lane, true, row_ok, ooff, n0, bias_s, nn[j],
                         reinterpret_cast<float*>(smem));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int BN>
int launch(const gb_conv_params& p, const CUtensorMap& ma, const CUtensorMap& mb, PairGeom pg, cudaStream_t st) {
  constexpr int B_BYTES = BN * BK * 2;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(igemm_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_set = true;
  }
  // smem plan: tail 2 KB (alignment slack + barriers + taps); A double-buffered when at least 3 weight stages
  // still fit, else single; the rest goes to weight stages (TMA latency ~1 us: bytes in flight matter)
  const int budget = SMEM_LIMIT - 2048;
  const int a_stage = NSUB * pg.a_sub_stride;
  int a_stages = (budget - 2 * a_stage) / B_BYTES >= 3 ? 2 : 1;
  if (g_gb_knobs[13] == 1 || g_gb_knobs[13] == 2) a_stages = g_gb_knobs[13];
  int total_chunks = pg.ngroups * (p.in.C >> 6);
  if (a_stages > total_chunks) a_stages = total_chunks;
  int b_stages = (budget - a_stages * a_stage) / B_BYTES;
  if (b_stages > MAX_B_STAGES) b_stages = MAX_B_STAGES;
  if (b_stages < 2) return -1;
  pg.a_stages = a_stages;
  pg.b_stages = b_stages;
  const int smem = a_stages * a_stage + b_stages * B_BYTES + 2048;
  dim3 grid(gb_cdiv(pg.nsub, NSUB), gb_cdiv(p.ncols, BN), 1);
  gb_klaunch(igemm_pair_kernel<BN>, grid, 256, smem, st, p, ma, mb, pg);
  g_gb_knobs[15] = 4;  // read-back slot: which data kernel served the last gb_conv_data call (tests)
  GB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int gb_tma_weight_map(const void* w, int kpad, int rows, int bn, CUtensorMap* out);  // igemm_tma.cu

// -1: not applicable (the caller tries the per-tap TMA kernel next), 0: launched, >0: error.
// knob 9: 1 = never use this kernel, 2 = use it whenever it applies (tests), 0 = its window (exactly one wave of 256-wide
// pair CTAs: the residual-block forward at batch 8).  Timed alone from a CUDA graph the per-tap kernel with the staged
// bulk-store epilogue is the faster one there (35.3 vs 39.1 us, profiles/r02u_conv_microbench_graph_b8.txt), but inside
// the two-stream step the 128-CTA launch leaves 20 SMs to the other chain and the step is 0.7 % faster with it
// (r02v: 396.1 / 400.1 vs 393.3 / 397.6 img/s), so the window stays.  knob 10: 1 = halo pitch 16 instead of kw + 7;
// knob 11: force the tile width; knob 13: force the number of activation stages.
int gb_conv_data_pair(const gb_conv_params& p, cudaStream_t st) {
  if (g_gb_knobs[9] == 1 || g_gb_knobs[3] != 0 || p.in_c_valid != 0) return -1;
  if (p.in.C % 64 != 0 || p.in.pad != 0 || !gb_tma_available()) return -1;
  for (int d = 0; d < 3; ++d)
    if (p.in_mul[d] != 1) return -1;
  if (p.nclass != 1 || p.ncols < 64) return -1;
  const gb_conv_class& cc = p.cls[0];
  if (cc.ntaps < 2 || cc.ntaps > GB_MAX_TAPS || cc.w_offset != 0) return -1;
  if (cc.ntaps * p.in.C > cc.kpad) return -1;
  if ((p.in.sx * 2) % 16 || (p.in.sy * 2) % 16 || (p.in.sz * 2) % 16 || (p.in.sn * 2) % 16) return -1;
  PairGeom pg;
  memset(&pg, 0, sizeof(pg));
  int dy_min = 127, dy_max = -128, dx_min = 127, dx_max = -128;
  for (int t = 0; t < cc.ntaps; ++t) {
    const int8_t* tp = p.taps[cc.tap_begin + t];
    dy_min = tp[1] < dy_min ? tp[1] : dy_min;
    dy_max = tp[1] > dy_max ? tp[1] : dy_max;
    dx_min = tp[2] < dx_min ? tp[2] : dx_min;
    dx_max = tp[2] > dx_max ? tp[2] : dx_max;
  }
  const int kh = dy_max - dy_min + 1, kw = dx_max - dx_min + 1;
  if (kw + TW - 1 > 16 || kh + TH - 1 > 24) return -1;
  // taps must be sorted by dz (itertools.product order); build the dz groups
  int ng = 0;
  for (int t = 0; t < cc.ntaps; ++t) {
    const int dz = p.taps[cc.tap_begin + t][0];
    if (ng == 0 || dz != pg.group_dz[ng - 1]) {
      if (ng >= 16) return -1;
      for (int g = 0; g < ng; ++g)
        if (pg.group_dz[g] == dz) return -1;
      pg.group_dz[ng] = (int8_t)dz;
      pg.group_begin[ng] = (int16_t)t;
      ++ng;
    }
  }
  pg.group_begin[ng] = (int16_t)cc.ntaps;
  pg.ngroups = ng;
  pg.dy_min = dy_min;
  pg.dx_min = dx_min;
  pg.hw = g_gb_knobs[10] == 1 ? 16 : kw + TW - 1;
  pg.hh = TH + kh - 1;
  pg.a_sub_bytes = pg.hw * pg.hh * 128;
  pg.a_sub_stride = (pg.a_sub_bytes + 1023) / 1024 * 1024;
  int q[3];
  gb_class_extents(p, 0, q);
  if (q[0] == 0 || q[1] == 0 || q[2] == 0 || p.in.N == 0) return 0;
  const int ntx = gb_cdiv(q[2], TW), nty = gb_cdiv(q[1], TH);
  pg.tiles_x = gb_make_fastdiv((uint32_t)ntx);
  pg.tiles_y = gb_make_fastdiv((uint32_t)nty);
  pg.tiles_z = gb_make_fastdiv((uint32_t)q[0]);
  const int64_t nsub = (int64_t)ntx * nty * q[0] * p.in.N;
  if (nsub >= (1ll << 30)) return -1;
  pg.nsub = (int)nsub;
  const int64_t npairs = (nsub + NSUB - 1) / NSUB;
  // tile width.  One CTA per SM; modelled time = rounds(148 SMs) x (MMA cycles of a CTA + fixed cost).  The kernel
  // only pays off when the launch fills most of the machine (small launches -- batch 1 -- are latency / occupancy
  // bound and keep the narrow-tile per-tap kernel, which spreads them over more SMs).
  int bn_max = 64;
  while (bn_max < p.ncols && bn_max < 256) bn_max *= 2;
  while (bn_max > p.npad) bn_max /= 2;
  if (bn_max < 64) return -1;
  int bn = 0;
  if (g_gb_knobs[11] > 0) {
    bn = g_gb_knobs[11];
    if (bn > bn_max) bn = bn_max;
  } else {
    bn = bn_max;
  }
  if (g_gb_knobs[9] != 2) {
    // Measured (profiles/r01f_conv_microbench_b8.txt): both this kernel and the per-tap kernel are bound by the
    // shared-memory operand reads of the single-CTA MMA (A 4 KB + B 8 KB per 128x256x16 instruction, ~54-60 %
    // tensor-pipe active), not by L2, so sharing B and re-using the halo buys little, and with two patches per
    // CTA a launch that is not ONE full wave loses more to wave quantisation than it gains.  Default: only the
    // launches that are exactly one wave of 256-wide pair CTAs on patch grids without waste (the residual-block
    // forward convolutions at batch 8: 128 CTAs, 45.9 us vs 48.1 us).
    const int64_t ctas = npairs * gb_cdiv(p.ncols, bn);
    const bool fits = (int64_t)ntx * TW * nty * TH * 100 <= (int64_t)q[2] * q[1] * 110;
    if (bn != 256 || ctas < 96 || ctas > 148 || !fits) return -1;
  }
  CUtensorMap ma, mb;
  if (gb_tma_activation_map(p.in, pg.hw, pg.hh, &ma)) return 1;
  if (gb_tma_weight_map(p.wpacked, cc.kpad, p.npad, bn, &mb)) return 1;
  switch (bn) {
    case 64: return launch<64>(p, ma, mb, pg, st);
    case 128: return launch<128>(p, ma, mb, pg, st);
    case 256: return launch<256>(p, ma, mb, pg, st);
  }
  return -1;
}
