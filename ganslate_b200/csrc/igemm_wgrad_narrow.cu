// Weight gradient of convolutions over NARROW tensors (16 or 32 channels per pixel on both sides, many taps): the
// 5x5x5 layers of the V-Net generators at the two finest resolutions -- the launches igemm_wgrad.cu served worst (its
// 128-row tiles are 3/4 padding for 32 output channels and the gathered operand travelled once per tap).
//
//   dW[(dz, dy, dx)][cin][cout] = sum_p  x[p + (dz, dy, dx)][cin] * dOut[p][cout]
//
// Output-stationary over the taps of ONE depth offset dz; the CTA walks 16 x 8 pixel patches.  Per patch it loads one
// halo box of x and one (x-extended) box of dOut, both with the swizzle whose span is a pixel, and reads them MN-major:
// the reduction index of the MMA is the pixel (two 8-pixel row segments per K = 16).
//   * M (rows of A) = `dy taps x Cin`: the atoms of a tile are consecutive ROWS of the halo (the descriptor's
//     leading-dimension stride is one halo row), 4 (Cin = 32) or 8 (Cin = 16) dy taps per 128-row tile.
//   * N (columns of B) = `dx taps x Cout`: the atoms are dOut shifted by 0, 1, ... kw - 1 PIXELS (leading-dimension
//     stride one pixel, so the atoms overlap in shared memory) against x read at the right-most tap:
//       D[(dy, ci)][(j, co)] = sum_p' x[p' + (dz, dy, dx_max)][ci] * dOut[p' + j][co] = dW[(dz, dy, dx_max - j)][ci][co].
//     Every (x pixel, dOut pixel) pair must meet in exactly one patch, so the patch grid starts kw - 1 pixels left of
//     the image (TMA zero-fills what lies outside).
// One MMA is 128 x (kw * Cout) x 16: with N = 160 / 80 columns the instruction is paced by the tensor math (or nearly),
// not by its shared-memory A fetch as N = 16 / 32 instructions are (~55 cycles whatever N; the first version of this
// kernel put the dx taps in M and issued 10 such MMAs per K step where this one issues 2).  No data moves per tap; x is
// read once per depth offset (5x) instead of once per tap (125x).  Rows of dy taps beyond kh accumulate garbage that
// the epilogue drops.
// Epilogue: TMEM -> registers -> red.global.add.f32 into the fp32 workspace dw[cout][tap * Cin + cin] (a warp's 32 lanes
// are 32 consecutive floats of one row); the pixel range is split over blockIdx.y.
#include <cuda.h>
#include <string.h>
#include "gb_common.cuh"
#include "gb_geometry.h"
#include "gb_tma.h"

namespace {

constexpr int TW = 8, TH = 16, HW = 16;  // patch 16 rows x 8 pixels; both boxes have a pitch of 16 pixels
constexpr int HH = TH + 7;               // halo rows: a tile's atoms reach 7 rows below the patch row
constexpr int MAX_KW = 8;                // dOut shifts 0 .. kw - 1 stay inside the 16-pixel pitch
constexpr int STAGES = 4;

template <int CG, int CP>
struct WNCfg {
  static constexpr int RBG = CG * 2, RBP = CP * 2;         // bytes per pixel
  static constexpr int APT = 128 / CG;                     // dy taps (atoms) per accumulator tile: 4 / 8
  static constexpr int A_BYTES = HW * HH * RBG;
  static constexpr int A_STRIDE = (A_BYTES + 1023) / 1024 * 1024;
  static constexpr int P_BYTES = TH * HW * RBP;            // multiple of 1 KB
  static constexpr int STAGE_BYTES = A_STRIDE + P_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 1024;
  static constexpr uint64_t LAYOUT_G = CG == 32 ? 4 : 6;   // SWIZZLE_64B / SWIZZLE_32B
  static constexpr uint64_t LAYOUT_P = CP == 32 ? 4 : 6;
};

struct WNGeom {
  gb_fastdiv tiles_x, tiles_y, tiles_z;
  int npatches, per_split;
  int kh, kw;
  int dy_min, dx_min, dx_max;
  int ntile;        // accumulator tiles: ceil(kh / APT)
  int ncol;         // columns of a tile: kw * Cout
  int ngroups;
  int8_t group_dz[16];
};

// MN-major operand: rows of the swizzle atom are pixels (`LAYOUT`'s span = one pixel); lbo = byte stride between
// atoms along M / N, sbo = byte stride between 8-pixel groups of the reduction.
template <uint64_t LAYOUT>
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= LAYOUT << 61;
  return d;
}

template <int CG, int CP>
__global__ void __launch_bounds__(256, 1)
igemm_wgrad_narrow_kernel(const __grid_constant__ gb_wgrad_params p, const __grid_constant__ CUtensorMap map_g,
                          const __grid_constant__ CUtensorMap map_p, const __grid_constant__ WNGeom wg) {
  gb_pdl_enter();
  using C = WNCfg<CG, CP>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* tail = smem + STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 128);
  int16_t* lut = reinterpret_cast<int16_t*>(tail + 192);   // [16 dy][8 dx]: tap index in this depth group, -1 = none

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int grp = blockIdx.x;
  const int dz = wg.group_dz[grp];
  const int b0 = blockIdx.y * wg.per_split;
  const int b1 = min(wg.npatches, b0 + wg.per_split);
  const int KB = b1 - b0;
  if (KB <= 0) return;

  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = smem_u32(bars + STAGES);
  const uint32_t accum_bar = smem_u32(bars + 2 * STAGES);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_slot));
  for (int i = tid; i < 16 * 8; i += 256) lut[i] = -1;
  __syncthreads();
  for (int t = tid; t < p.ntaps; t += 256)
    if (p.taps[t][0] == dz) lut[(p.taps[t][1] - wg.dy_min) * 8 + (p.taps[t][2] - wg.dx_min)] = (int16_t)t;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (gb_elect_one()) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES, it = kb / STAGES;
        if (it > 0) mbar_wait(empty_bar + 8 * s, (it - 1) & 1);
        uint32_t t = (uint32_t)(b0 + kb);
        uint32_t u = gb_div(t, wg.tiles_x);
        const int x0 = (int)(t - u * wg.tiles_x.d) * TW - (wg.kw - 1);   // the patch grid starts left of the image
        t = u;
        u = gb_div(t, wg.tiles_y);
        const int y0 = (int)(t - u * wg.tiles_y.d) * TH;
        t = u;
        u = gb_div(t, wg.tiles_z);
        const int z0 = (int)(t - u * wg.tiles_z.d);
        const int n = (int)u;
        const uint32_t a_s = base + s * C::STAGE_BYTES;
        const uint32_t bar = full_bar + 8 * s;
        mbar_expect_tx(bar, (uint32_t)(C::A_BYTES + C::P_BYTES));
        // x at the right-most tap: box pixel i <-> x = x0 + dx_max + i;  dOut: box pixel i <-> x = x0 + i
        tma_load_5d(a_s, &map_g, bar, 0, x0 + wg.dx_max, y0 + wg.dy_min, z0 + dz, n);
        tma_load_5d(a_s + C::A_STRIDE, &map_p, bar, 0, x0, y0, z0, n);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // One thread issues every MMA; the region is guarded by elect.sync so that the descriptor arithmetic stays in the
    // uniform datapath (gb_elect_one).
    if (gb_elect_one()) {
      const uint32_t idesc = make_idesc_bf16(wg.ncol, 1, 1);
      // A: atoms = dy taps (one halo row apart), 8-pixel groups = patch rows (one halo row apart)
      const uint64_t adesc0 = make_desc_mn<C::LAYOUT_G>(base, HW * C::RBG, HW * C::RBG);
      // B: atoms = dOut shifted by one more pixel each, 8-pixel groups = patch rows
      const uint64_t bdesc0 = make_desc_mn<C::LAYOUT_P>(base + C::A_STRIDE, C::RBP, HW * C::RBP);
      constexpr uint64_t A_ROW = (HW * C::RBG) >> 4;          // one halo row (descriptor units of 16 bytes)
      constexpr uint64_t B_ROW = (HW * C::RBP) >> 4;
      const int ntile = wg.ntile;
      const uint32_t ncol = (uint32_t)wg.ncol;
      uint32_t accumulate = 0;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES, it = kb / STAGES;
        mbar_wait(full_bar + 8 * s, it & 1);
        tc_fence_after();
        const uint64_t a_st = adesc0 + (uint64_t)((s * C::STAGE_BYTES) >> 4);
        const uint64_t b_st = bdesc0 + (uint64_t)((s * C::STAGE_BYTES) >> 4);
#pragma unroll 1
        for (int ks = 0; ks < TH / 2; ++ks) {   // K = 16 pixels: rows 2 ks and 2 ks + 1 of the patch
          const uint64_t bdesc = b_st + (uint64_t)(2 * ks) * B_ROW;
          uint64_t adesc = a_st + (uint64_t)(2 * ks) * A_ROW;
          uint32_t tcol = tmem_base;
          for (int ti = 0; ti < ntile; ++ti, adesc += C::APT * A_ROW, tcol += ncol)
            umma_bf16(tcol, adesc, bdesc, idesc, accumulate);
          accumulate = 1;
        }
        umma_commit(empty_bar + 8 * s);
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }

  mbar_wait(accum_bar, 0);
  tc_fence_after();
  {
    const int lg = warp & 3;
    const int m = lg * 32 + lane;           // accumulator row: dy atom a of the tile, input channel c
    const int a = m / CG, c = m - a * CG;
    // work items (tile, dx shift j): the two warp groups take alternate items
    const int nitem = wg.ntile * wg.kw;
    for (int item = (warp >> 2); item < nitem; item += 2) {
      const int ti = item / wg.kw, j = item - ti * wg.kw;
      const int dyi = ti * C::APT + a;
      const int tl = dyi < 16 ? lut[dyi * 8 + (wg.kw - 1 - j)] : -1;   // column block j <-> dx = dx_max - j
      uint32_t acc[CP];
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ti * wg.ncol + j * CP);
      if constexpr (CP == 32) tmem_ld32(taddr, acc);
      else tmem_ld16(taddr, acc);
      tmem_ld_wait();
      if (tl >= 0) {
        float* dst = p.dw + (int64_t)tl * CG + c;
#pragma unroll
        for (int nn = 0; nn < CP; ++nn)
          if (nn < p.rows) atomicAdd(dst + (int64_t)nn * p.kpad, __uint_as_float(acc[nn]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

template <int CG, int CP>
int launch(const gb_wgrad_params& p, const CUtensorMap& mg, const CUtensorMap& mp, const WNGeom& wg, int splits,
           cudaStream_t st) {
  using C = WNCfg<CG, CP>;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(igemm_wgrad_narrow_kernel<CG, CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  gb_klaunch(igemm_wgrad_narrow_kernel<CG, CP>, dim3(wg.ngroups, splits, 1), 256, C::SMEM, st, p, mg, mp, wg);
  g_gb_knobs[14] = 3;  // read-back slot: which wgrad variant served the last call (tests)
  GB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int gb_tma_activation_map_narrow(const gb_view& v, int cbox, int tw, int th, CUtensorMap* out);  // igemm_tma.cu

// -1: not applicable, 0: launched, > 0: error.  knob 12 = 5 switches this variant off.
int gb_conv_wgrad_narrow(const gb_wgrad_params& p, cudaStream_t st) {
  if (g_gb_knobs[12] == 5 || g_gb_knobs[3] != 0 || !gb_tma_available() || p.gathered_c_valid != 0) return -1;
  const int cg = p.gathered.C, cp = p.plain.C;
  if (!(cg == 16 || cg == 32) || !(cp == 16 || cp == 32)) return -1;
  if (p.mul[0] != 1 || p.mul[1] != 1 || p.mul[2] != 1 || p.plain.pad != 0 || p.gathered.pad != 0) return -1;
  if (p.ntaps < 9 || p.rows > cp) return -1;
  for (const gb_view* v : {&p.plain, &p.gathered})
    if (((uintptr_t)v->ptr & 15) != 0 || (v->sx * 2) % 16 != 0 || (v->sy * 2) % 16 != 0 || (v->sz * 2) % 16 != 0 ||
        (v->sn * 2) % 16 != 0)
      return -1;
  WNGeom wg;
  memset(&wg, 0, sizeof(wg));
  int dy_min = 127, dy_max = -128, dx_min = 127, dx_max = -128;
  for (int t = 0; t < p.ntaps; ++t) {
    dy_min = p.taps[t][1] < dy_min ? p.taps[t][1] : dy_min;
    dy_max = p.taps[t][1] > dy_max ? p.taps[t][1] : dy_max;
    dx_min = p.taps[t][2] < dx_min ? p.taps[t][2] : dx_min;
    dx_max = p.taps[t][2] > dx_max ? p.taps[t][2] : dx_max;
    const int dzv = p.taps[t][0];
    bool seen = false;
    for (int g = 0; g < wg.ngroups; ++g) seen = seen || wg.group_dz[g] == dzv;
    if (!seen) {
      if (wg.ngroups >= 16) return -1;
      wg.group_dz[wg.ngroups++] = (int8_t)dzv;
    }
  }
  wg.kh = dy_max - dy_min + 1;
  wg.kw = dx_max - dx_min + 1;
  const int apt = 128 / cg;
  wg.ntile = gb_cdiv(wg.kh, apt);
  wg.ncol = wg.kw * cp;
  // dOut shifts inside the 16-pixel pitch; one MMA's N <= 256; the accumulators inside the TMEM; every atom of a tile
  // inside the halo box (garbage rows included)
  if (wg.kw > MAX_KW || wg.ncol > 256 || wg.ncol % 16 != 0 || wg.ntile * wg.ncol > 512 || wg.ntile * apt > 8) return -1;
  wg.dy_min = dy_min;
  wg.dx_min = dx_min;
  wg.dx_max = dx_max;
  const int ntx = gb_cdiv(p.plain.W + wg.kw - 1, TW), nty = gb_cdiv(p.plain.H, TH);
  if ((int64_t)ntx * TW * nty * TH * 100 > (int64_t)p.plain.W * p.plain.H * 175) return -1;  // patches must fit the image
  const int64_t np = (int64_t)ntx * nty * p.plain.D * p.plain.N;
  if (np <= 0 || np >= (1ll << 31)) return -1;
  wg.tiles_x = gb_make_fastdiv((uint32_t)ntx);
  wg.tiles_y = gb_make_fastdiv((uint32_t)nty);
  wg.tiles_z = gb_make_fastdiv((uint32_t)p.plain.D);
  wg.npatches = (int)np;
  // one CTA per SM (the accumulators take most of the TMEM): one wave, at least two patches per CTA
  int splits = p.splits > 0 ? p.splits : 148 / wg.ngroups;
  if (splits > np / 2) splits = (int)(np / 2);
  if (splits < 1) splits = 1;
  wg.per_split = gb_cdiv(np, splits);
  splits = gb_cdiv(np, wg.per_split);
  CUtensorMap mg, mp;
  if (gb_tma_activation_map_narrow(p.gathered, cg, HW, HH, &mg)) return 1;
  if (gb_tma_activation_map_narrow(p.plain, cp, HW, TH, &mp)) return 1;
  if (cg == 32) return cp == 32 ? launch<32, 32>(p, mg, mp, wg, splits, st) : launch<32, 16>(p, mg, mp, wg, splits, st);
  return cp == 32 ? launch<16, 32>(p, mg, mp, wg, splits, st) : launch<16, 16>(p, mg, mp, wg, splits, st);
}
