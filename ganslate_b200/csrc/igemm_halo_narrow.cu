// Halo-reuse implicit-GEMM convolution for NARROW inputs: 16 or 32 channels per pixel, many taps -- the 5x5x5 layers of
// the V-Net generators at the two finest resolutions (ganslate/nn/generators/vnet/vnet3d.py: 16 -> 16 and 32 -> 32 on
// 32 x 256 x 256 voxels are 60 % of a generator pass's FLOPs), forward and data gradient.
//
// A 64-channel K block of igemm_tma / igemm_halo does not exist here: a pixel is 32 or 64 bytes.  The operand layouts
// follow the pixel instead:
//   * A: per depth offset dz ONE halo box {C ch, 16 px, TH + kh - 1 rows} lands in shared memory with the swizzle whose
//     span is the pixel (SWIZZLE_32B / SWIZZLE_64B), 8-pixel groups one image row (16 x C x 2 bytes) apart.  The A tile
//     of tap (dz, dy, dx) is that image read through a K-major descriptor shifted by (dy * 16 + dx) pixels: 25 taps per
//     box, no data movement per tap.  `tcgen05.mma` K = 16 channels: one MMA per tap for C = 16, two for C = 32.
//   * B: the packed weights keep their K index (tap, channel) and their 64-wide SWIZZLE_128B boxes; a box holds 64 / C
//     consecutive taps and tap j of a box starts j * C * 2 bytes into the 128-byte rows.  A and B descriptors carry
//     different swizzle modes; boxes may straddle two depth groups (125 taps do not divide), so the weight ring is
//     decoupled from the halo ring and both are walked in tap order.
// Per 128 output pixels the per-tap kernel moved 125 x 128 pixels through L2 -> smem (and the cp.async gather issued
// four 16-byte copies per pixel and tap); here it is 5 boxes of 16 x 20 pixels (6 % of that) plus the weights.
// With N = 16 / 32 columns the MMA is bound by its shared-memory A read (4 KB per 128 x N x 16 instruction), not by the
// tensor pipe: about a quarter of the dense peak is the ceiling for these layer shapes.
#include <cuda.h>
#include "gb_common.cuh"
#include "gb_geometry.h"
#include "gb_epilogue.cuh"
#include "gb_tma.h"

namespace {

constexpr int TW = 8, TH = 16, HW = 16;  // patch 16 x 8 output pixels, halo pitch 16 pixels
constexpr int MAX_HH = 24;               // halo rows (TH + kh - 1), kh <= 9
constexpr int A_STAGES = 2;
constexpr int B_STAGES = 3;

template <int BN, int CIN>
struct NCfg {
  static constexpr int RB = CIN * 2;                      // bytes per pixel row
  static constexpr int TPB = 64 / CIN;                    // taps per 64-wide weight box
  static constexpr int KSTEPS = CIN / 16;                 // MMAs per tap
  static constexpr int BLK_BYTES = BN * 128;              // one weight box
  static constexpr int TG = (16 * 1024 / BLK_BYTES) > 8 ? 8 : (16 * 1024 / BLK_BYTES);  // boxes per B stage
  static constexpr int BS_BYTES = TG * BLK_BYTES;
  static constexpr int A_BYTES_MAX = HW * MAX_HH * RB;    // 24 KB (C = 32) / 12 KB (C = 16): multiples of 1 KB
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM = A_STAGES * A_BYTES_MAX + B_STAGES * BS_BYTES + 1024 + 1024;
  static constexpr uint64_t LAYOUT = CIN == 32 ? 4 : 6;   // UMMA layout type: SWIZZLE_64B / SWIZZLE_32B
};

struct NarrowGeom {
  gb_fastdiv tiles_x, tiles_y, tiles_z;
  int ntiles;
  int hh;           // halo rows
  int a_bytes;      // HW * hh * RB
  int ngroups;      // distinct dz values (1 for 2-D)
  int dy_min, dx_min;
  int8_t group_dz[16];
  int16_t group_begin[17];  // taps are sorted by dz: taps [group_begin[g], group_begin[g+1]) share dz
};

// K-major operand whose swizzle span is one pixel row; 8-row groups `sbo_bytes` apart.  The swizzle XOR is taken from
// the absolute shared-memory address (measured for SWIZZLE_128B in igemm_halo.cu), so a start address shifted by whole
// pixels reads the image TMA wrote.
template <uint64_t LAYOUT>
__device__ __forceinline__ uint64_t make_desc_px(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                              // LBO (unused: K = 16 elements stay inside the swizzle span)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= LAYOUT << 61;
  return d;
}

template <int BN, int CIN>
__global__ void __launch_bounds__(256, 2)
igemm_halo_narrow_kernel(const __grid_constant__ gb_conv_params p, const __grid_constant__ CUtensorMap map_a,
                         const __grid_constant__ CUtensorMap map_b, const __grid_constant__ NarrowGeom hg) {
  gb_pdl_enter();
  using C = NCfg<BN, CIN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base;
  const uint32_t b_base = base + A_STAGES * C::A_BYTES_MAX;
  uint8_t* tail = smem + A_STAGES * C::A_BYTES_MAX + B_STAGES * C::BS_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  // layout: a_full[A_STAGES], a_empty[A_STAGES], b_full[B_STAGES], b_empty[B_STAGES], accum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 256);
  int8_t* taps_s = reinterpret_cast<int8_t*>(tail + 320);
  __shared__ float bias_s[BN];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const gb_conv_class& cc = p.cls[0];
  int q[3];
  gb_class_extents(p, 0, q);
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
  const int ntaps = cc.ntaps;
  const int nblk = (ntaps + C::TPB - 1) / C::TPB;        // weight boxes
  const int nbst = (nblk + C::TG - 1) / C::TG;           // B stages
  constexpr int TAPS_PER_STAGE = C::TG * C::TPB;

  const uint32_t a_full = smem_u32(bars), a_empty = smem_u32(bars + A_STAGES);
  const uint32_t b_full = smem_u32(bars + 2 * A_STAGES), b_empty = smem_u32(bars + 2 * A_STAGES + B_STAGES);
  const uint32_t accum_bar = smem_u32(bars + 2 * A_STAGES + 2 * B_STAGES);
  if (tid == 0) {
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(a_full + 8 * s, 1);
      mbar_init(a_empty + 8 * s, 1);
    }
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(smem_u32(tmem_slot));
  for (int i = tid; i < ntaps; i += 256)
    *reinterpret_cast<uint32_t*>(taps_s + 4 * i) = *reinterpret_cast<const uint32_t*>(p.taps[cc.tap_begin + i]);
  for (int i = tid; i < BN; i += 256) bias_s[i] = (p.bias != nullptr && n0 + i < p.ncols) ? p.bias[n0 + i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      auto load_a = [&](int g) {
        const int as = g % A_STAGES, it = g / A_STAGES;
        if (it > 0) mbar_wait(a_empty + 8 * as, (it - 1) & 1);
        mbar_expect_tx(a_full + 8 * as, (uint32_t)hg.a_bytes);
        tma_load_5d(a_base + as * C::A_BYTES_MAX, &map_a, a_full + 8 * as, 0, x0 + hg.dx_min, y0 + hg.dy_min,
                    z0 + hg.group_dz[g], n);
      };
      // halo boxes run one depth group ahead of the weights: group g + 1 is requested when the stage holding the first
      // tap of group g goes out (every weight stage the release of its buffer depends on has been issued by then)
      int next_a = 0;
      for (; next_a < hg.ngroups && next_a < A_STAGES; ++next_a) load_a(next_a);
      int g_cur = 0;
      for (int s = 0; s < nbst; ++s) {
        const int bs = s % B_STAGES, it = s / B_STAGES;
        const int blk0 = s * C::TG, nb = min(C::TG, nblk - blk0);
        if (it > 0) mbar_wait(b_empty + 8 * bs, (it - 1) & 1);
        mbar_expect_tx(b_full + 8 * bs, (uint32_t)(nb * C::BLK_BYTES));
        for (int j = 0; j < nb; ++j)
          tma_load_2d(b_base + bs * C::BS_BYTES + j * C::BLK_BYTES, &map_b, b_full + 8 * bs, (blk0 + j) * 64, n0);
        const int last_tap = min(ntaps, (s + 1) * TAPS_PER_STAGE) - 1;
        while (g_cur + 1 < hg.ngroups && hg.group_begin[g_cur + 1] <= last_tap) {
          ++g_cur;   // this stage carries the first tap of group g_cur
          if (next_a < hg.ngroups && next_a <= g_cur + 1) load_a(next_a++);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(BN, 0, 0);
    uint32_t first = 1;
    int g = 0;
    for (int tp = 0; tp < ntaps; ++tp) {
      while (tp >= hg.group_begin[g + 1]) ++g;
      const int as = g % A_STAGES;
      if (tp == hg.group_begin[g]) mbar_wait(a_full + 8 * as, (g / A_STAGES) & 1);
      const int s = tp / TAPS_PER_STAGE, bs = s % B_STAGES;
      if (tp == s * TAPS_PER_STAGE) mbar_wait(b_full + 8 * bs, (s / B_STAGES) & 1);
      tc_fence_after();
      const bool last_of_group = tp + 1 == hg.group_begin[g + 1];
      const bool last_of_stage = tp + 1 == ntaps || (tp + 1) % TAPS_PER_STAGE == 0;
      if (lane == 0) {
        const int ry = taps_s[4 * tp + 1] - hg.dy_min, rx = taps_s[4 * tp + 2] - hg.dx_min;
        const uint32_t a_s = a_base + as * C::A_BYTES_MAX + (uint32_t)(ry * HW + rx) * C::RB;
        const uint64_t adesc = make_desc_px<C::LAYOUT>(a_s, HW * C::RB);
        const int jb = tp - s * TAPS_PER_STAGE;   // tap within the stage: box jb / TPB, position jb % TPB in its rows
        const uint64_t bdesc = make_smem_desc(b_base + bs * C::BS_BYTES + (jb / C::TPB) * C::BLK_BYTES, 16, 1024) +
                               (uint64_t)((jb % C::TPB) * C::RB / 16);
#pragma unroll
        for (int k = 0; k < C::KSTEPS; ++k) {
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, first ? 0u : 1u);
          first = 0;
        }
        if (last_of_stage) umma_commit(b_empty + 8 * bs);
        if (last_of_group) umma_commit(a_empty + 8 * as);
      }
      __syncwarp();
    }
    if (lane == 0 && ntaps > 0) umma_commit(accum_bar);
    __syncwarp();
  }

  if (ntaps > 0) {
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
    gb_conv_epilogue<BN>(p, tmem_base, warp, lane, ntaps > 0, row_ok, ooff, n0, bias_s, n, reinterpret_cast<float*>(smem));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int BN, int CIN>
int launch(const gb_conv_params& p, const CUtensorMap& ma, const CUtensorMap& mb, const NarrowGeom& hg, cudaStream_t st) {
  using C = NCfg<BN, CIN>;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(igemm_halo_narrow_kernel<BN, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  dim3 grid(hg.ntiles, gb_cdiv(p.ncols, BN), 1);
  gb_klaunch(igemm_halo_narrow_kernel<BN, CIN>, grid, 256, C::SMEM, st, p, ma, mb, hg);
  g_gb_knobs[15] = 8;  // read-back slot: which data kernel served the last gb_conv_data call (tests)
  GB_LAUNCH_CHECK();
  return 0;
}

template <int CIN>
int launch_bn(int bn, const gb_conv_params& p, const CUtensorMap& ma, const CUtensorMap& mb, const NarrowGeom& hg,
              cudaStream_t st) {
  switch (bn) {
    case 16: return launch<16, CIN>(p, ma, mb, hg, st);
    case 32: return launch<32, CIN>(p, ma, mb, hg, st);
    case 64: return launch<64, CIN>(p, ma, mb, hg, st);
  }
  return -1;
}

}  // namespace

int gb_tma_weight_map(const void* w, int kpad, int rows, int bn, CUtensorMap* out);  // igemm_tma.cu
// igemm_tma.cu: 5-D map {C, W, H, D, N} with box {cbox ch, tw, th, 1, 1}, cbox = 16 / 32 -> SWIZZLE_32B / SWIZZLE_64B
int gb_tma_activation_map_narrow(const gb_view& v, int cbox, int tw, int th, CUtensorMap* out);

// -1: not applicable, 0: launched, >0 error.  Unit-stride single-class convolutions over exactly 16 or 32 input channels
// with at least 16 taps and at most 64 output columns; knob 4 = 2 switches every halo kernel off, knob 4 = 3 this one only.
int gb_conv_data_halo_narrow(const gb_conv_params& p, cudaStream_t st) {
  if (g_gb_knobs[4] == 2 || g_gb_knobs[4] == 3 || g_gb_knobs[3] != 0 || p.in_c_valid != 0) return -1;
  if (!(p.in.C == 16 || p.in.C == 32) || p.in.pad != 0 || !gb_tma_available()) return -1;
  if (p.nclass != 1 || p.ncols > 64) return -1;
  for (int d = 0; d < 3; ++d)
    if (p.in_mul[d] != 1) return -1;
  const gb_conv_class& cc = p.cls[0];
  if (cc.ntaps < 16 || cc.ntaps > GB_MAX_TAPS || cc.w_offset != 0) return -1;
  // the activation view as a tensor map: 16-byte aligned base and strides
  if (((uintptr_t)p.in.ptr & 15) != 0 || (p.in.sx * 2) % 16 != 0 || (p.in.sy * 2) % 16 != 0 || (p.in.sz * 2) % 16 != 0 ||
      (p.in.sn * 2) % 16 != 0)
    return -1;
  NarrowGeom hg;
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
  int ng = 0;
  for (int t = 0; t < cc.ntaps; ++t) {  // taps must be sorted by dz (they are: itertools.product order)
    const int dz = p.taps[cc.tap_begin + t][0];
    if (ng == 0 || dz != hg.group_dz[ng - 1]) {
      if (ng >= 16) return -1;
      for (int g = 0; g < ng; ++g)
        if (hg.group_dz[g] == dz) return -1;
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
  hg.a_bytes = HW * hg.hh * p.in.C * 2;
  int q[3];
  gb_class_extents(p, 0, q);
  if (q[0] == 0 || q[1] == 0 || q[2] == 0) return 0;
  const int ntx = gb_cdiv(q[2], TW), nty = gb_cdiv(q[1], TH);
  if ((int64_t)ntx * TW * nty * TH * 100 > (int64_t)q[2] * q[1] * 135) return -1;  // patches must fit the image
  hg.tiles_x = gb_make_fastdiv((uint32_t)ntx);
  hg.tiles_y = gb_make_fastdiv((uint32_t)nty);
  hg.tiles_z = gb_make_fastdiv((uint32_t)q[0]);
  const int64_t ntiles = (int64_t)ntx * nty * q[0] * p.in.N;
  if (ntiles >= (1ll << 31)) return -1;
  hg.ntiles = (int)ntiles;
  int bn = 16;
  while (bn < p.ncols && bn < 64) bn *= 2;
  if (bn > p.npad) return -1;
  CUtensorMap ma, mb;
  if (gb_tma_activation_map_narrow(p.in, p.in.C, HW, hg.hh, &ma)) return 1;
  if (gb_tma_weight_map(p.wpacked, cc.kpad, p.npad, bn, &mb)) return 1;
  return p.in.C == 32 ? launch_bn<32>(bn, p, ma, mb, hg, st) : launch_bn<16>(bn, p, ma, mb, hg, st);
}
