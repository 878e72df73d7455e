// Halo-reuse variant of the TMA implicit-GEMM convolution.
//
// igemm_tma.cu loads one A tile per (tap, 64-channel chunk): the same input pixels travel L2 -> smem once per tap
// (9x for a 3x3, 49x for a 7x7 convolution).  Here a CTA owns a 16-row x 8-column patch of output pixels and loads,
// per 64-channel chunk, ONE halo box {64 ch, 16 px, 16+kh-1 rows} (pitch 16 pixels = 2 KB).  Because the patch is
// 8 pixels wide, every 8-row swizzle group of the A operand is one image row of the halo, groups are a constant
// 2 KB apart, and the A tile of tap (ry, rx) is simply the same smem image read through a descriptor whose start
// address is shifted by (ry*16 + rx) rows -- no data movement per tap.  Weights stream per (tap, chunk) as before.
//
// A traffic drops by kh*kw*128/(16*(16+kh-1)) (4.0x for 3x3, 17.8x for 7x7); for the 7x7 layers with 3 or 64
// output channels (BN = 16/64), which are pure im2col-traffic bound in igemm_tma.cu, that is the whole runtime.
#include <cuda.h>
#include "gb_common.cuh"
#include "gb_geometry.h"
#include "gb_epilogue.cuh"
#include "gb_tma.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int TW = 8, TH = 16, HW = 16;  // patch 16 x 8 output pixels, halo pitch 16 pixels
constexpr int MAX_HH = 24;               // halo rows (TH + kh - 1), kh <= 9

template <int BN>
struct HCfg {
  static constexpr int B_BYTES = BN * BK * 2;            // one tap x 64 channels of weights
  // narrow tiles (the 7x7 -> 3 channel layers) are bound by the barrier round trip per tap, not by data: a B stage
  // carries TG taps (one wait / commit per TG*4 MMAs) and two CTAs share an SM so that one CTA's halo load and
  // epilogue overlap the other's MMAs
  static constexpr int TG = (BN <= 32) ? (16 * 1024 / B_BYTES > 8 ? 8 : 16 * 1024 / B_BYTES) : 1;
  static constexpr int A_BYTES_MAX = HW * MAX_HH * 128;  // 48 KB
  static constexpr int A_STAGES = (BN <= 32) ? 1 : 2;
  static constexpr int B_STAGES = (BN <= 32) ? 3 : ((BN == 256) ? 3 : (BN == 128 ? 6 : 8));
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM = A_STAGES * A_BYTES_MAX + B_STAGES * TG * B_BYTES + 1024 + 1024;
  static constexpr int MIN_CTAS = (BN <= 32) ? 2 : 1;
};

struct HaloGeom {
  gb_fastdiv tiles_x, tiles_y, tiles_z;
  int ntiles;
  int hh;           // halo rows
  int a_bytes;      // HW * hh * 128
  int ngroups;      // distinct dz values (1 for 2-D)
  int dy_min, dx_min;
  int8_t group_dz[16];
  int16_t group_begin[17];  // taps are sorted by dz: taps [group_begin[g], group_begin[g+1]) share dz
};

__device__ __forceinline__ uint64_t make_smem_desc_bo(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                              // LBO (unused for swizzled K-major)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  // matrix base offset: measured on B200 -- the swizzle XOR uses the absolute smem address bits [7,10), so a start
  // address that is only 128 B aligned needs base_offset 0 (setting (addr >> 7) & 7 gives wrong results)
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN>
__global__ void __launch_bounds__(256, HCfg<BN>::MIN_CTAS)
igemm_halo_kernel(const __grid_constant__ gb_conv_params p, const __grid_constant__ CUtensorMap map_a,
                  const __grid_constant__ CUtensorMap map_b, const __grid_constant__ HaloGeom hg, int base_offset_mode) {
  gb_pdl_enter();
  using C = HCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base;
  constexpr int TG = C::TG;
  constexpr int BS_BYTES = TG * C::B_BYTES;  // bytes of one B stage
  const uint32_t b_base = base + C::A_STAGES * C::A_BYTES_MAX;
  uint8_t* tail = smem + C::A_STAGES * C::A_BYTES_MAX + C::B_STAGES * BS_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  // layout: a_full[2], a_empty[2], b_full[B_STAGES], b_empty[B_STAGES], accum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 256);
  int8_t* taps_s = reinterpret_cast<int8_t*>(tail + 320);
  __shared__ float bias_s[BN];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cls = blockIdx.z;
  const gb_conv_class& cc = p.cls[cls];
  int q[3];
  gb_class_extents(p, cls, q);
  uint32_t t = blockIdx.x;
  uint32_t u = gb_div(t, hg.tiles_x);
  const int tx = (int)(t - u * hg.tiles_x.d);
  t = u;
  u = gb_div(t, hg.tiles_y);
  const int ty = (int)(t - u * hg.tiles_y.d);
  t = u;
  u = gb_div(t, hg.tiles_z);
  const int z0 = (int)(t - u * hg.tiles_z.d);
  const int n = (int)u;
  const int x0 = tx * TW, y0 = ty * TH;
  if (n >= p.in.N || z0 >= q[0] || y0 >= q[1] || x0 >= q[2]) return;
  const int n0 = blockIdx.y * BN;
  const int chunks = p.in.C >> 6;

  const uint32_t a_full = smem_u32(bars), a_empty = smem_u32(bars + 2);
  const uint32_t b_full = smem_u32(bars + 4), b_empty = smem_u32(bars + 4 + C::B_STAGES);
  const uint32_t accum_bar = smem_u32(bars + 4 + 2 * C::B_STAGES);
  if (tid == 0) {
    for (int s = 0; s < C::A_STAGES; ++s) {
      mbar_init(a_full + 8 * s, 1);
      mbar_init(a_empty + 8 * s, 1);
    }
    for (int s = 0; s < C::B_STAGES; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(smem_u32(tmem_slot));
  for (int i = tid; i < cc.ntaps; i += 256)
    *reinterpret_cast<uint32_t*>(taps_s + 4 * i) = *reinterpret_cast<const uint32_t*>(p.taps[cc.tap_begin + i]);
  for (int i = tid; i < BN; i += 256) bias_s[i] = (p.bias != nullptr && n0 + i < p.ncols) ? p.bias[n0 + i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_mma_groups = hg.ngroups * chunks;

  if (warp == 0) {
    if (gb_elect_one()) {
      int ai = 0, bi = 0;
      for (int g = 0; g < hg.ngroups; ++g) {
        const int dz = hg.group_dz[g];
        for (int c = 0; c < chunks; ++c, ++ai) {
          const int as = ai % C::A_STAGES, ait = ai / C::A_STAGES;
          if (ait > 0) mbar_wait(a_empty + 8 * as, (ait - 1) & 1);
          mbar_expect_tx(a_full + 8 * as, (uint32_t)hg.a_bytes);
          tma_load_5d(a_base + as * C::A_BYTES_MAX, &map_a, a_full + 8 * as, c * 64, x0 + hg.dx_min, y0 + hg.dy_min,
                      z0 + dz, n);
          for (int tl = hg.group_begin[g]; tl < hg.group_begin[g + 1]; tl += TG, ++bi) {
            const int bs = bi % C::B_STAGES, bit = bi / C::B_STAGES;
            const int nt = min(TG, hg.group_begin[g + 1] - tl);
            if (bit > 0) mbar_wait(b_empty + 8 * bs, (bit - 1) & 1);
            mbar_expect_tx(b_full + 8 * bs, (uint32_t)(nt * C::B_BYTES));
            for (int j = 0; j < nt; ++j)
              tma_load_2d(b_base + bs * BS_BYTES + j * C::B_BYTES, &map_b, b_full + 8 * bs, (tl + j) * p.in.C + c * 64,
                          cls * p.npad + n0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(BN, 0, 0);
    int ai = 0, bi = 0;
    uint32_t first = 1;
    for (int g = 0; g < hg.ngroups; ++g) {
      for (int c = 0; c < chunks; ++c, ++ai) {
        const int as = ai % C::A_STAGES, ait = ai / C::A_STAGES;
        mbar_wait(a_full + 8 * as, ait & 1);
        for (int tl = hg.group_begin[g]; tl < hg.group_begin[g + 1]; tl += TG, ++bi) {
          const int bs = bi % C::B_STAGES, bit = bi / C::B_STAGES;
          const int nt = min(TG, hg.group_begin[g + 1] - tl);
          mbar_wait(b_full + 8 * bs, bit & 1);
          tc_fence_after();
          if (gb_elect_one()) {
            for (int j = 0; j < nt; ++j) {
              const int ry = taps_s[4 * (tl + j) + 1] - hg.dy_min, rx = taps_s[4 * (tl + j) + 2] - hg.dx_min;
              const uint32_t a_s = a_base + as * C::A_BYTES_MAX + (uint32_t)(ry * HW + rx) * 128u;
              const uint32_t bo = base_offset_mode ? ((a_s >> 7) & 7u) : 0u;
              const uint64_t adesc = make_smem_desc_bo(a_s, HW * 128, bo);
              const uint64_t bdesc = make_smem_desc(b_base + bs * BS_BYTES + j * C::B_BYTES, 16, 1024);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, first ? 0u : 1u);
                first = 0;
              }
            }
            umma_commit(b_empty + 8 * bs);
          }
          __syncwarp();
        }
        if (lane == 0) umma_commit(a_empty + 8 * as);
        __syncwarp();
      }
    }
    if (lane == 0 && total_mma_groups > 0) umma_commit(accum_bar);
    __syncwarp();
  }

  if (total_mma_groups > 0 && cc.ntaps > 0) {
    mbar_wait(accum_bar, 0);
    tc_fence_after();
  }
  {
    const int row = (warp & 3) * 32 + lane;
    const int h = row >> 3, w = row & 7;
    const int qy = y0 + h, qx = x0 + w;
    const bool row_ok = qy < q[1] && qx < q[2];
    int64_t ooff = 0;
    if (row_ok)
      ooff = gb_pix_offset(p.out, n, z0 * p.out_mul[0] + cc.off[0], qy * p.out_mul[1] + cc.off[1],
                           qx * p.out_mul[2] + cc.off[2]);
    gb_conv_epilogue<BN>(p, tmem_base, warp, lane, cc.ntaps > 0, row_ok, ooff, n0, bias_s, n, reinterpret_cast<float*>(smem));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int BN>
int launch(const gb_conv_params& p, const CUtensorMap& ma, const CUtensorMap& mb, const HaloGeom& hg, cudaStream_t st) {
  using C = HCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(igemm_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  dim3 grid(hg.ntiles, gb_cdiv(p.ncols, BN), p.nclass);
  gb_klaunch(igemm_halo_kernel<BN>, grid, 256, C::SMEM, st, p, ma, mb, hg, g_gb_knobs[5] == 2 ? 1 : 0);
  g_gb_knobs[15] = 3;
  GB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int gb_tma_weight_map(const void* w, int kpad, int rows, int bn, CUtensorMap* out);  // igemm_tma.cu

// -1: not applicable, 0: launched, >0 error.
// Used by default only for narrow outputs with many taps (the generators' 7x7 convolutions to / from 3 channels:
// BN <= 32), where the per-tap kernel re-reads the same 64-channel pixels 49 times from L2 for an MMA of 8 cycles
// (measured: 424 us at batch 8 for a layer whose tensors stream in 20 us).  For wide tiles the first version of
// this kernel measured 5-15 % slower than the per-tap kernel (one barrier round trip per tap, one CTA per SM), so
// those stay on igemm_tma unless knob 4 = 1 forces this path; knob 4 = 2 disables it.
int gb_conv_data_halo(const gb_conv_params& p, cudaStream_t st) {
  if (g_gb_knobs[4] == 2 || g_gb_knobs[3] != 0 || p.in_c_valid != 0) return -1;
  if (g_gb_knobs[4] != 1 && !(p.ncols <= 32 && p.nclass == 1 && p.cls[0].ntaps >= 16)) return -1;
  if (p.in.C % 64 != 0 || p.in.pad != 0 || !gb_tma_available()) return -1;
  for (int d = 0; d < 3; ++d)
    if (p.in_mul[d] != 1) return -1;
  if (p.nclass != 1) return -1;  // (parity classes have different windows; they stay on igemm_tma for now)
  const gb_conv_class& cc = p.cls[0];
  if (cc.ntaps < 2 || cc.ntaps > GB_MAX_TAPS) return -1;
  if (cc.w_offset != 0) return -1;
  HaloGeom hg;
  memset(&hg, 0, sizeof(hg));
  int dy_min = 127, dy_max = -128, dx_min = 127, dx_max = -128;
  for (int t = 0; t < cc.ntaps; ++t) {
    const int8_t* tp = p.taps[cc.tap_begin + t];
    dy_min = tp[1] < dy_min ? tp[1] : dy_min;
    dy_max = tp[1] > dy_max ? tp[1] : dy_max;
    dx_min = tp[2] < dx_min ? tp[2] : dx_min;
    dx_max = tp[2] > dx_max ? tp[2] : dx_max;
  }
  const int kh = dy_max - dy_min + 1, kw = dx_max - dx_min + 1;
  if (kw + TW - 1 > HW || kh + TH - 1 > MAX_HH) return -1;
  // taps must be sorted by dz (they are: itertools.product order); build the dz groups
  int ng = 0;
  for (int t = 0; t < cc.ntaps; ++t) {
    const int dz = p.taps[cc.tap_begin + t][0];
    if (ng == 0 || dz != hg.group_dz[ng - 1]) {
      if (ng >= 16) return -1;
      for (int g = 0; g < ng; ++g)
        if (hg.group_dz[g] == dz) return -1;  // not sorted by dz
      hg.group_dz[ng] = (int8_t)dz;
      hg.group_begin[ng] = (int16_t)t;
      ++ng;
    }
  }
  hg.group_begin[ng] = (int16_t)cc.ntaps;
  hg.ngroups = ng;
  hg.dy_min = dy_min;
  hg.dx_min = dx_min;
  hg.hh = TH + kh - 1;
  hg.a_bytes = HW * hg.hh * 128;
  int q[3];
  gb_class_extents(p, 0, q);
  if (q[0] == 0 || q[1] == 0 || q[2] == 0) return 0;
  const int ntx = gb_cdiv(q[2], TW), nty = gb_cdiv(q[1], TH);
  // only worth it when the 16x8 patches fit the image reasonably (waste < 35 %)
  if ((int64_t)ntx * TW * nty * TH * 100 > (int64_t)q[2] * q[1] * 135) return -1;
  hg.tiles_x = gb_make_fastdiv((uint32_t)ntx);
  hg.tiles_y = gb_make_fastdiv((uint32_t)nty);
  hg.tiles_z = gb_make_fastdiv((uint32_t)q[0]);
  const int64_t ntiles = (int64_t)ntx * nty * q[0] * p.in.N;
  if (ntiles >= (1ll << 31)) return -1;
  hg.ntiles = (int)ntiles;
  int bn = 16;
  while (bn < p.ncols && bn < 256) bn *= 2;
  if (g_gb_knobs[1] > 0) bn = g_gb_knobs[1];
  else
    while (bn > 64 && ntiles * gb_cdiv(p.ncols, bn) < 148) bn /= 2;
  if (bn > p.npad) return -1;
  CUtensorMap ma, mb;
  if (gb_tma_activation_map(p.in, HW, hg.hh, &ma)) return 1;
  if (gb_tma_weight_map(p.wpacked, cc.kpad, p.npad, bn, &mb)) return 1;
  switch (bn) {
    case 16: return launch<16>(p, ma, mb, hg, st);
    case 32: return launch<32>(p, ma, mb, hg, st);
    case 64: return launch<64>(p, ma, mb, hg, st);
    case 128: return launch<128>(p, ma, mb, hg, st);
    case 256: return launch<256>(p, ma, mb, hg, st);
  }
  return -1;
}
