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
// With N = 16 / 32 columns the MMA is bound by its shared-memory A fetch, not by the tensor math: ~55 cycles per
// 128 x N x 16 instruction whatever N (measured), i.e. ~600 TFLOP/s is the ceiling of this formulation for 32 -> 32
// layers and ~340 for 16 -> 16 (the per-tap gather kernel it replaces ran them at 240 / 110).
#include <cuda.h>
#include "gb_common.cuh"
#include "gb_geometry.h"
#include "gb_epilogue.cuh"
#include "gb_tma.h"

namespace {

constexpr int TW = 8, TH = 16, HW = 16;  // patch 16 x 8 output pixels, halo pitch 16 pixels
constexpr int MAX_KH = 9;                // halo rows = NP * TH + kh - 1

// NP = patches per CTA, stacked in y (NP accumulators of BN columns).  One MMA of N = 16 / 32 columns lasts ~17 cycles,
// so a CTA is bound by the ONE thread that issues them and by the weights it streams (every tap's box once per CTA):
// with four patches a tap is 4 x KSTEPS instructions behind one table read and one descriptor update, the weights
// travel once per 512 pixels, and the halo box (NP * 16 + kh - 1 rows) carries less overlap.  NP = 1 serves images
// that four-patch columns do not fit.
template <int BN, int CIN, int NP>
struct NCfg {
  static constexpr int RB = CIN * 2;                      // bytes per pixel row
  static constexpr int TPB = 64 / CIN;                    // taps per 64-wide weight box
  static constexpr int KSTEPS = CIN / 16;                 // MMAs per tap and patch
  static constexpr int BLK_BYTES = BN * 128;              // one weight box
  static constexpr int STAGE_TARGET = (CIN == 16 && NP > 1) ? 8 * 1024 : 16 * 1024;
  static constexpr int TG = (STAGE_TARGET / BLK_BYTES) > 8 ? 8 : (STAGE_TARGET / BLK_BYTES);  // boxes per B stage
  static constexpr int BS_BYTES = TG * BLK_BYTES;
  // Two CTAs per SM in every configuration: an MMA of 128 x (16 | 32) x 16 occupies the tensor pipe for ~55 cycles
  // whatever N is (its shared-memory A fetch; measured: 1000 MMAs per 512-pixel tile take 63 K cycles in a lone CTA with
  // deep rings, 52 K per CTA when two CTAs interleave), so what a second CTA adds is overlap of the other's prologue,
  // epilogue and ring bubbles.  32 channels x four patches: single-buffered 68 KB halo, two weight stages.
  static constexpr int MIN_CTAS = 2;
  static constexpr int A_STAGES = (NP > 1 && CIN == 32) ? 1 : 2;
  static constexpr int B_STAGES = (NP > 1 && CIN == 32) ? 2 : ((NP > 1) ? 4 : 3);
  static constexpr int A_BYTES_MAX = HW * (NP * TH + MAX_KH - 1) * RB;   // multiple of 1 KB for NP = 1, 4
  static constexpr int PATCH_BYTES = TH * HW * RB;        // halo rows of one patch
  static constexpr int TMEM_COLS = NP * BN < 32 ? 32 : NP * BN;
  static constexpr int SMEM = A_STAGES * A_BYTES_MAX + B_STAGES * BS_BYTES + 1024 + 1024;
  static constexpr uint64_t LAYOUT = CIN == 32 ? 4 : 6;   // UMMA layout type: SWIZZLE_64B / SWIZZLE_32B
  static_assert(A_BYTES_MAX % 1024 == 0, "halo stages must keep the swizzle phase");
  static_assert(MIN_CTAS * SMEM <= 227 * 1024, "shared memory per SM");
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

template <int BN, int CIN, int NP>
__global__ void __launch_bounds__(256, NCfg<BN, CIN, NP>::MIN_CTAS)
igemm_halo_narrow_kernel(const __grid_constant__ gb_conv_params p, const __grid_constant__ CUtensorMap map_a,
                         const __grid_constant__ CUtensorMap map_b, const __grid_constant__ NarrowGeom hg) {
  gb_pdl_enter();
  using C = NCfg<BN, CIN, NP>;
  constexpr int A_STAGES = C::A_STAGES, B_STAGES = C::B_STAGES;
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
  uint16_t* a_off_s = reinterpret_cast<uint16_t*>(tail + 320);   // per tap: (ry * HW + rx) * RB / 16
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
  const int x0 = tx * TW, y0 = ty * (NP * TH);
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
  for (int i = tid; i < ntaps; i += 256) {
    const int8_t* tp = p.taps[cc.tap_begin + i];
    a_off_s[i] = (uint16_t)((((int)tp[1] - hg.dy_min) * HW + ((int)tp[2] - hg.dx_min)) * C::RB / 16);
  }
  for (int i = tid; i < BN; i += 256) bias_s[i] = (p.bias != nullptr && n0 + i < p.ncols) ? p.bias[n0 + i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (gb_elect_one()) {
      auto load_a = [&](int g) {
        const int as = g % A_STAGES, it = g / A_STAGES;
        if (it > 0) mbar_wait(a_empty + 8 * as, (it - 1) & 1);
        mbar_expect_tx(a_full + 8 * as, (uint32_t)hg.a_bytes);
        tma_load_5d(a_base + as * C::A_BYTES_MAX, &map_a, a_full + 8 * as, 0, x0 + hg.dx_min, y0 + hg.dy_min,
                    z0 + hg.group_dz[g], n);
      };
      // A halo box may be requested once every weight stage that the release of its buffer depends on has gone out:
      // box g reuses the buffer of group g - A_STAGES, which is released after that group's last tap.
      int next_a = 0;
      for (; next_a < hg.ngroups && next_a < A_STAGES; ++next_a) load_a(next_a);
      for (int s = 0; s < nbst; ++s) {
        const int bs = s % B_STAGES, it = s / B_STAGES;
        const int blk0 = s * C::TG, nb = min(C::TG, nblk - blk0);
        if (it > 0) mbar_wait(b_empty + 8 * bs, (it - 1) & 1);
        mbar_expect_tx(b_full + 8 * bs, (uint32_t)(nb * C::BLK_BYTES));
        for (int j = 0; j < nb; ++j)
          tma_load_2d(b_base + bs * C::BS_BYTES + j * C::BLK_BYTES, &map_b, b_full + 8 * bs, (blk0 + j) * 64, n0);
        const int last_tap = min(ntaps, (s + 1) * TAPS_PER_STAGE) - 1;
        while (next_a < hg.ngroups && hg.group_begin[next_a - A_STAGES + 1] - 1 <= last_tap) load_a(next_a++);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // One thread issues every MMA, and an MMA of N = 16 / 32 columns lasts ~32 cycles: the loop body is a table read
    // (prefetched one tap ahead), two additions and the instruction itself; barrier waits only at stage / group starts.
    if (ntaps > 0 && gb_elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BN, 0, 0);
      const uint64_t adesc0 = make_desc_px<C::LAYOUT>(a_base, HW * C::RB);
      const uint64_t bdesc0 = make_smem_desc(b_base, 16, 1024);
      uint32_t accumulate = 0;
      int g = 0, g_end = hg.group_begin[1];
      uint64_t adesc_g = adesc0;
      mbar_wait(a_full, 0);
      uint32_t a_off = a_off_s[0];
      int tp = 0;
      for (int s = 0; s < nbst; ++s) {
        const int bs = s % B_STAGES;
        mbar_wait(b_full + 8 * bs, (s / B_STAGES) & 1);
        tc_fence_after();
        const uint64_t bdesc_s = bdesc0 + (uint64_t)((bs * C::BS_BYTES) >> 4);
        const int tap_end = min(ntaps, (s + 1) * TAPS_PER_STAGE);
        for (int jb = 0; tp < tap_end; ++tp, ++jb) {
          if (tp == g_end) {   // first tap of the next depth group: its halo box
            ++g;
            g_end = hg.group_begin[g + 1];
            const int as = g % A_STAGES;
            mbar_wait(a_full + 8 * as, (g / A_STAGES) & 1);
            tc_fence_after();
            adesc_g = adesc0 + (uint64_t)((as * C::A_BYTES_MAX) >> 4);
          }
          const uint64_t adesc = adesc_g + a_off;
          a_off = a_off_s[min(tp + 1, ntaps - 1)];
          const uint64_t bdesc = bdesc_s + (uint64_t)((jb / C::TPB) * (C::BLK_BYTES >> 4) + (jb % C::TPB) * (C::RB >> 4));
#pragma unroll
          for (int pi = 0; pi < NP; ++pi)
#pragma unroll
            for (int k = 0; k < C::KSTEPS; ++k)
              umma_bf16(tmem_base + (uint32_t)(pi * BN), adesc + (uint64_t)(pi * (C::PATCH_BYTES >> 4) + 2 * k), bdesc + 2 * k,
                        idesc, (k == 0) ? accumulate : 1u);
          accumulate = 1;
          if (tp + 1 == g_end) umma_commit(a_empty + 8 * (g % A_STAGES));
        }
        umma_commit(b_empty + 8 * bs);
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }

  if (ntaps > 0) {
    mbar_wait(accum_bar, 0);
    tc_fence_after();
  }
  {
    const int row = (warp & 3) * 32 + lane;
    const int h = row >> 3, w = row & 7;
    const int qx = x0 + w;
    // BN < 64: the shared epilogue uses four warps per accumulator, so the two warp groups take alternate patches, and
    // the InstanceNorm statistics of all patches are summed in shared memory first (the idle halo buffer)
    constexpr bool SPLIT = BN < 64;
    float* sstats = reinterpret_cast<float*>(smem);
    const bool smem_stats = SPLIT && p.stats != nullptr && !p.out_fp32;
    if (smem_stats) {
      for (int i = tid; i < 2 * BN; i += 256) sstats[i] = 0.f;
      __syncthreads();
    }
    for (int pi = SPLIT ? (warp >> 2) : 0; pi < NP; pi += SPLIT ? 2 : 1) {
      const int qy = y0 + pi * TH + h;
      const bool row_ok = qy < q[1] && qx < q[2];
      int64_t ooff = 0;
      if (row_ok)
        ooff = gb_pix_offset(p.out, n, z0 * p.out_mul[0] + cc.off[0], qy * p.out_mul[1] + cc.off[1],
                             qx * p.out_mul[2] + cc.off[2]);
      gb_conv_epilogue<BN>(p, tmem_base + (uint32_t)(pi * BN), SPLIT ? (warp & 3) : warp, lane, ntaps > 0, row_ok, ooff, n0,
                           bias_s, n, SPLIT ? nullptr : sstats, smem_stats ? sstats : nullptr);
    }
    if (smem_stats) {
      __syncthreads();
      for (int i = tid; i < 2 * BN; i += 256)
        if (n0 + (i >> 1) < p.ncols) atomicAdd(p.stats + ((int64_t)n * p.out.C + n0 + (i >> 1)) * 2 + (i & 1), sstats[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int BN, int CIN, int NP>
int launch(const gb_conv_params& p, const CUtensorMap& ma, const CUtensorMap& mb, const NarrowGeom& hg, cudaStream_t st) {
  using C = NCfg<BN, CIN, NP>;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(igemm_halo_narrow_kernel<BN, CIN, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  dim3 grid(hg.ntiles, gb_cdiv(p.ncols, BN), 1);
  gb_klaunch(igemm_halo_narrow_kernel<BN, CIN, NP>, grid, 256, C::SMEM, st, p, ma, mb, hg);
  g_gb_knobs[15] = 8;  // read-back slot: which data kernel served the last gb_conv_data call (tests)
  GB_LAUNCH_CHECK();
  return 0;
}

template <int CIN, int NP>
int launch_bn(int bn, const gb_conv_params& p, const CUtensorMap& ma, const CUtensorMap& mb, const NarrowGeom& hg,
              cudaStream_t st) {
  switch (bn) {
    case 16: return launch<16, CIN, NP>(p, ma, mb, hg, st);
    case 32: return launch<32, CIN, NP>(p, ma, mb, hg, st);
    case 64: return launch<64, CIN, NP>(p, ma, mb, hg, st);
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
  if (kw + TW - 1 > HW || kh > MAX_KH) return -1;
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
  int q[3];
  gb_class_extents(p, 0, q);
  if (q[0] == 0 || q[1] == 0 || q[2] == 0) return 0;
  // four patches per CTA where columns of 64 rows fit the image (knob 13 = 3: always one patch), else one
  const int ntx = gb_cdiv(q[2], TW);
  int np = 4;
  if (g_gb_knobs[13] == 3 || (int64_t)ntx * TW * gb_cdiv(q[1], 4 * TH) * 4 * TH * 100 > (int64_t)q[2] * q[1] * 135) np = 1;
  const int nty = gb_cdiv(q[1], np * TH);
  if ((int64_t)ntx * TW * nty * np * TH * 100 > (int64_t)q[2] * q[1] * 135) return -1;  // patches must fit the image
  hg.hh = np * TH + kh - 1;
  hg.a_bytes = HW * hg.hh * p.in.C * 2;
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
  if (np == 4) return p.in.C == 32 ? launch_bn<32, 4>(bn, p, ma, mb, hg, st) : launch_bn<16, 4>(bn, p, ma, mb, hg, st);
  return p.in.C == 32 ? launch_bn<32, 1>(bn, p, ma, mb, hg, st) : launch_bn<16, 1>(bn, p, ma, mb, hg, st);
}
