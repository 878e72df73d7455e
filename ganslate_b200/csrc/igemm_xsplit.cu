// x-split implicit-GEMM convolution for NARROW layers: 16 or 32 input channels, at most 32 output channels, many taps --
// forward and data gradient of the 5x5x5 V-Net layers at the two finest resolutions.
//
// igemm_halo_narrow.cu issues one `tcgen05.mma` of 128 x Cout x 16 per tap (and 16 channels): with N = 16 / 32 columns
// such an instruction is paced by its shared-memory A fetch, ~55 cycles whatever N, a quarter of the tensor rate.  The
// dx taps of a kernel row all read the SAME activation rows, shifted by one pixel each.  So the shift is moved from
// the operand to the result:
//
//   P_j[p'][co] = sum_{dz, dy, c} x[p' + (dz, dy, 0)][c] * W[(dz, dy, dx_j)][co][c]          (all j in ONE GEMM)
//   out[p][co]  = sum_j P_j[p + dx_j][co]                                                     (epilogue)
//
// The accumulator tile is 128 pixels (8 rows x 16 columns of the halo box, which is contiguous in shared memory: no
// shifted descriptors at all) x N = kw * Cout columns (160 for 5 x 32): one MMA per (dz, dy) and 16 channels instead
// of kw, paced by the tensor math.  The weights of a kernel row land as kw boxes {C, Cout} stacked along N.  The
// epilogue reads the kw column blocks of a row from TMEM and adds them with warp shuffles (block j comes from the lane
// j pixels to the right); of the 16 columns of a tile row 16 - (kw - 1) are outputs.
// A CTA owns NP such tiles stacked in y (NP * N <= 512 TMEM columns) so that the weights travel once per NP tiles; one
// stage of the ring = one depth offset (halo box + kh * kw weight boxes), two stages.
#include <cuda.h>
#include <string.h>
#include "gb_common.cuh"
#include "gb_geometry.h"
#include "gb_epilogue.cuh"
#include "gb_tma.h"

namespace {

constexpr int TY = 8, HW = 16;          // tile: 8 rows x 16 halo columns
constexpr int MAX_KH = 8, MAX_KW = 8, MAX_G = 8;
constexpr int STAGES = 2;

template <int CIN, int CO>
struct XCfg {
  static constexpr int RB = CIN * 2;                      // bytes per pixel / per weight row
  static constexpr int KSTEPS = CIN / 16;
  static constexpr int NP_MAX = 4;
  static constexpr int A_BYTES_MAX = HW * (NP_MAX * TY + MAX_KH - 1) * RB;     // 39 rows: 39 KB (C = 32) / 19.5 KB
  static constexpr int A_STRIDE = (A_BYTES_MAX + 1023) / 1024 * 1024;
  static constexpr int B_BYTES_MAX = 5 * 5 * CO * RB;     // kernel rows x columns budget: kh * kw <= 25 boxes
  static constexpr int B_STRIDE = (B_BYTES_MAX + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = A_STRIDE + B_STRIDE;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 1024;
  static constexpr uint64_t LAYOUT = CIN == 32 ? 4 : 6;   // SWIZZLE_64B / SWIZZLE_32B
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

struct XGeom {
  gb_fastdiv tiles_x, tiles_y, tiles_z;
  int ntiles;
  int np;           // tiles per CTA (stacked in y)
  int kh, kw, xw;   // xw = output columns per tile row = 16 - (kw - 1)
  int hh;           // halo rows = np * TY + kh - 1
  int a_bytes, b_bytes;
  int ncol;         // kw * CO
  int dy_min, dx_min;
  int ngroups;
  int8_t group_dz[MAX_G];
  int16_t tap[MAX_G][MAX_KH][MAX_KW];   // tap index of (group, dy, dx)
};

// K-major operand with rows of one pixel (`LAYOUT`'s span), 8-row groups contiguous
template <uint64_t LAYOUT>
__device__ __forceinline__ uint64_t make_desc_rows(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= LAYOUT << 61;
  return d;
}

template <int CIN, int CO>
__global__ void __launch_bounds__(256, 1)
igemm_xsplit_kernel(const __grid_constant__ gb_conv_params p, const __grid_constant__ CUtensorMap map_a,
                    const __grid_constant__ CUtensorMap map_b, const __grid_constant__ XGeom xg) {
  gb_pdl_enter();
  using C = XCfg<CIN, CO>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* tail = smem + STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 128);
  __shared__ float bias_s[CO];
  __shared__ float sstats[2 * CO];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const gb_conv_class& cc = p.cls[0];
  int q[3];
  gb_class_extents(p, 0, q);
  uint32_t t = blockIdx.x;
  uint32_t u = gb_div(t, xg.tiles_x);
  const int tx = (int)(t - u * xg.tiles_x.d);
  t = u;
  u = gb_div(t, xg.tiles_y);
  const int ty = (int)(t - u * xg.tiles_y.d);
  t = u;
  u = gb_div(t, xg.tiles_z);
  const int z0 = (int)(t - u * xg.tiles_z.d);
  const int n = (int)u;
  const int x0 = tx * xg.xw, y0 = ty * (xg.np * TY);
  if (n >= p.in.N || z0 >= q[0] || y0 >= q[1] || x0 >= q[2]) return;

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
  for (int i = tid; i < CO; i += 256) bias_s[i] = (p.bias != nullptr && i < p.ncols) ? p.bias[i] : 0.f;
  for (int i = tid; i < 2 * CO; i += 256) sstats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ng = xg.ngroups, kh = xg.kh, kw = xg.kw, np = xg.np;

  if (warp == 0) {
    if (gb_elect_one()) {
      for (int g = 0; g < ng; ++g) {
        const int s = g % STAGES, it = g / STAGES;
        if (it > 0) mbar_wait(empty_bar + 8 * s, (it - 1) & 1);
        const uint32_t a_s = base + s * C::STAGE_BYTES;
        const uint32_t b_s = a_s + C::A_STRIDE;
        const uint32_t bar = full_bar + 8 * s;
        mbar_expect_tx(bar, (uint32_t)(xg.a_bytes + xg.b_bytes));
        tma_load_5d(a_s, &map_a, bar, 0, x0 + xg.dx_min, y0 + xg.dy_min, z0 + xg.group_dz[g], n);
        for (int dyi = 0; dyi < kh; ++dyi)
          for (int j = 0; j < kw; ++j)
            tma_load_2d(b_s + (uint32_t)((dyi * kw + j) * CO) * C::RB, &map_b, bar, (int)xg.tap[g][dyi][j] * CIN, 0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (gb_elect_one()) {
      const uint32_t idesc = make_idesc_bf16(xg.ncol, 0, 0);
      const uint64_t adesc0 = make_desc_rows<C::LAYOUT>(base, 8 * C::RB);
      const uint64_t bdesc0 = make_desc_rows<C::LAYOUT>(base + C::A_STRIDE, 8 * C::RB);
      constexpr uint64_t A_ROW = (HW * C::RB) >> 4;               // one halo row, in descriptor units of 16 bytes
      const uint64_t b_row = (uint64_t)((kw * CO * C::RB) >> 4);  // the weights of one kernel row
      const uint32_t ncol = (uint32_t)xg.ncol;
      uint32_t accumulate = 0;
      for (int g = 0; g < ng; ++g) {
        const int s = g % STAGES;
        mbar_wait(full_bar + 8 * s, (g / STAGES) & 1);
        tc_fence_after();
        const uint64_t a_st = adesc0 + (uint64_t)((s * C::STAGE_BYTES) >> 4);
        uint64_t bdesc = bdesc0 + (uint64_t)((s * C::STAGE_BYTES) >> 4);
#pragma unroll 1
        for (int dyi = 0; dyi < kh; ++dyi, bdesc += b_row) {
          uint64_t adesc = a_st + (uint64_t)dyi * A_ROW;
          uint32_t tcol = tmem_base;
          for (int pi = 0; pi < np; ++pi, adesc += TY * A_ROW, tcol += ncol) {
#pragma unroll
            for (int k = 0; k < C::KSTEPS; ++k) umma_bf16(tcol, adesc + 2 * k, bdesc + 2 * k, idesc, (k == 0) ? accumulate : 1u);
          }
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
    // TMEM lane = tile row = y_local * 16 + x'; a warp's 32 lanes are two tile rows, so the shuffles stay inside it
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    const int yl = row >> 4, xp = row & 15;
    const int qx = x0 + xp;
    const bool want_stats = p.stats != nullptr && !p.out_fp32;
    __nv_bfloat16* optr = reinterpret_cast<__nv_bfloat16*>(p.out.ptr);
    for (int pi = (warp >> 2); pi < np; pi += 2) {
      float v[CO];
#pragma unroll
      for (int i = 0; i < CO; ++i) v[i] = bias_s[i];
      for (int j = 0; j < kw; ++j) {
        uint32_t acc[CO];
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(pi * xg.ncol + j * CO);
        if constexpr (CO == 32) tmem_ld32(taddr, acc);
        else tmem_ld16(taddr, acc);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < CO; ++i) v[i] += __shfl_down_sync(0xffffffffu, __uint_as_float(acc[i]), j);
      }
      const int qy = y0 + pi * TY + yl;
      const bool row_ok = xp < xg.xw && qy < q[1] && qx < q[2];
      int64_t ooff = 0;
      if (row_ok) ooff = gb_pix_offset(p.out, n, z0 + cc.off[0], qy + cc.off[1], qx + cc.off[2]);
#pragma unroll
      for (int i = 0; i < CO; ++i) {
        float tv = v[i];
        if (p.act == GB_ACT_TANH) tv = tanhf(tv);
        else if (p.act == GB_ACT_LEAKY) tv = tv > 0.f ? tv : tv * p.act_slope;
        else if (p.act == GB_ACT_RELU) tv = fmaxf(tv, 0.f);
        v[i] = tv;
      }
      if (p.out_fp32) {
        if (row_ok) {
          float* o = reinterpret_cast<float*>(p.out.ptr) + ooff;
#pragma unroll
          for (int g8 = 0; g8 < CO / 4; ++g8) {
            if (g8 * 4 < p.out.C) {
              float4* o4 = reinterpret_cast<float4*>(o + g8 * 4);
              float4 a = make_float4(v[g8 * 4], v[g8 * 4 + 1], v[g8 * 4 + 2], v[g8 * 4 + 3]);
              if (p.accumulate) {
                const float4 pa = *o4;
                a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
              }
              *o4 = a;
            }
          }
        }
      } else {
        float sv[CO];
#pragma unroll
        for (int g8 = 0; g8 < CO / 8; ++g8) {
          uint4 o;
          o.x = pack_bf16x2(v[g8 * 8 + 0], v[g8 * 8 + 1]);
          o.y = pack_bf16x2(v[g8 * 8 + 2], v[g8 * 8 + 3]);
          o.z = pack_bf16x2(v[g8 * 8 + 4], v[g8 * 8 + 5]);
          o.w = pack_bf16x2(v[g8 * 8 + 6], v[g8 * 8 + 7]);
          if (row_ok && g8 * 8 < p.out.C) *reinterpret_cast<uint4*>(optr + ooff + g8 * 8) = o;
          float2 f;
          f = unpack_bf16x2(o.x); sv[g8 * 8 + 0] = f.x; sv[g8 * 8 + 1] = f.y;
          f = unpack_bf16x2(o.y); sv[g8 * 8 + 2] = f.x; sv[g8 * 8 + 3] = f.y;
          f = unpack_bf16x2(o.z); sv[g8 * 8 + 4] = f.x; sv[g8 * 8 + 5] = f.y;
          f = unpack_bf16x2(o.w); sv[g8 * 8 + 6] = f.x; sv[g8 * 8 + 7] = f.y;
        }
        if (want_stats) {
          float sq[CO];
#pragma unroll
          for (int i = 0; i < CO; ++i) {
            if (!row_ok || i >= p.ncols) sv[i] = 0.f;
            sq[i] = sv[i] * sv[i];
          }
          const float s1 = gb_warp_colsum<CO>(sv, lane);
          const float s2 = gb_warp_colsum<CO>(sq, lane);
          if (lane < CO) {
            atomicAdd(sstats + lane * 2, s1);
            atomicAdd(sstats + lane * 2 + 1, s2);
          }
        }
      }
    }
    if (want_stats) {
      __syncthreads();
      for (int i = tid; i < 2 * CO; i += 256)
        if ((i >> 1) < p.ncols) atomicAdd(p.stats + ((int64_t)n * p.out.C + (i >> 1)) * 2 + (i & 1), sstats[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

template <int CIN, int CO>
int launch(const gb_conv_params& p, const CUtensorMap& ma, const CUtensorMap& mb, const XGeom& xg, cudaStream_t st) {
  using C = XCfg<CIN, CO>;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(igemm_xsplit_kernel<CIN, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  gb_klaunch(igemm_xsplit_kernel<CIN, CO>, dim3(xg.ntiles, 1, 1), 256, C::SMEM, st, p, ma, mb, xg);
  g_gb_knobs[15] = 9;  // read-back slot: which data kernel served the last gb_conv_data call (tests)
  GB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int gb_tma_activation_map_narrow(const gb_view& v, int cbox, int tw, int th, CUtensorMap* out);       // igemm_tma.cu
int gb_tma_weight_map_narrow(const void* w, int kpad, int rows, int cbox, int bn, CUtensorMap* out);  // igemm_tma.cu

// -1: not applicable, 0: launched, > 0: error.  Unit-stride single-class convolutions over exactly 16 or 32 input
// channels with a complete kd x kh x kw tap box (kh, kw <= 5 x 5 boxes per depth offset, kw >= 2) and 16 / 32 packed
// output rows.  knob 4 = 2 switches every halo kernel off, knob 4 = 4 this one only (igemm_halo_narrow then serves),
// knob 4 = 5 widens it from its default (16 -> 16 layers) to everything it supports.
int gb_conv_data_xsplit(const gb_conv_params& p, cudaStream_t st) {
  if (g_gb_knobs[4] == 2 || g_gb_knobs[4] == 3 || g_gb_knobs[4] == 4 || g_gb_knobs[3] != 0 || p.in_c_valid != 0) return -1;
  // Measured (profiles/r02ae_conv3d_microbench_xsplit.txt): an MMA's operand fetch grows with M + N rows (~125 cycles for
  // 128 x 160 x 16), and with the accumulators of three tiles in TMEM only one CTA fits an SM, so prologue and epilogue
  // are exposed.  16 -> 16 layers: 375 / 341 us against 404 / 392 us on igemm_halo_narrow (forward / data gradient on
  // 32 x 256 x 256 voxels); 32 -> 32 layers: 876 / 853 us against 722 / 768 us.  Default: the 16 -> 16 case only;
  // knob 4 = 5 takes every layer this kernel supports.
  if (g_gb_knobs[4] != 5 && !(p.in.C == 16 && p.npad == 16)) return -1;
  if (!(p.in.C == 16 || p.in.C == 32) || !(p.npad == 16 || p.npad == 32) || p.in.pad != 0 || !gb_tma_available()) return -1;
  if (p.nclass != 1 || p.ncols > p.npad) return -1;
  for (int d = 0; d < 3; ++d)
    if (p.in_mul[d] != 1 || p.out_mul[d] != 1) return -1;
  const gb_conv_class& cc = p.cls[0];
  if (cc.ntaps < 9 || cc.ntaps > GB_MAX_TAPS || cc.w_offset != 0) return -1;
  if (((uintptr_t)p.in.ptr & 15) != 0 || (p.in.sx * 2) % 16 != 0 || (p.in.sy * 2) % 16 != 0 || (p.in.sz * 2) % 16 != 0 ||
      (p.in.sn * 2) % 16 != 0)
    return -1;
  // the fp32 / bf16 stores of the epilogue are 16-byte vectors
  const int esz = p.out_fp32 ? 4 : 2;
  if (((uintptr_t)p.out.ptr & 15) != 0 || (p.out.sx * esz) % 16 != 0 || (p.out.sy * esz) % 16 != 0 ||
      (p.out.sz * esz) % 16 != 0 || (p.out.sn * esz) % 16 != 0 || p.out.C % 8 != 0)
    return -1;
  XGeom xg;
  memset(&xg, 0, sizeof(xg));
  int dy_min = 127, dy_max = -128, dx_min = 127, dx_max = -128;
  for (int t = 0; t < cc.ntaps; ++t) {
    const int8_t* tp = p.taps[cc.tap_begin + t];
    dy_min = tp[1] < dy_min ? tp[1] : dy_min;
    dy_max = tp[1] > dy_max ? tp[1] : dy_max;
    dx_min = tp[2] < dx_min ? tp[2] : dx_min;
    dx_max = tp[2] > dx_max ? tp[2] : dx_max;
    bool seen = false;
    for (int g = 0; g < xg.ngroups; ++g) seen = seen || xg.group_dz[g] == tp[0];
    if (!seen) {
      if (xg.ngroups >= MAX_G) return -1;
      xg.group_dz[xg.ngroups++] = tp[0];
    }
  }
  const int kh = dy_max - dy_min + 1, kw = dx_max - dx_min + 1;
  if (kh > MAX_KH || kw > MAX_KW || kw < 2 || kh * kw > 25 || cc.ntaps != xg.ngroups * kh * kw) return -1;
  for (int g = 0; g < xg.ngroups; ++g)
    for (int a = 0; a < kh; ++a)
      for (int b = 0; b < kw; ++b) xg.tap[g][a][b] = -1;
  for (int t = 0; t < cc.ntaps; ++t) {
    const int8_t* tp = p.taps[cc.tap_begin + t];
    int g = 0;
    while (xg.group_dz[g] != tp[0]) ++g;
    if (xg.tap[g][tp[1] - dy_min][tp[2] - dx_min] != -1) return -1;  // a repeated tap
    xg.tap[g][tp[1] - dy_min][tp[2] - dx_min] = (int16_t)t;
  }
  const int co = p.npad;
  xg.ncol = kw * co;
  if (xg.ncol > 256 || xg.ncol % 16 != 0) return -1;
  xg.kh = kh;
  xg.kw = kw;
  xg.xw = HW - (kw - 1);
  xg.dy_min = dy_min;
  xg.dx_min = dx_min;
  int q[3];
  gb_class_extents(p, 0, q);
  if (q[0] == 0 || q[1] == 0 || q[2] == 0) return 0;
  int np = 512 / xg.ncol;
  if (np > 4) np = 4;
  while (np > 1 && gb_cdiv(q[1], np * TY) * np * TY * 100 > q[1] * 125) --np;   // rows beyond the image are wasted MMAs
  const int ntx = gb_cdiv(q[2], xg.xw), nty = gb_cdiv(q[1], np * TY);
  if ((int64_t)ntx * xg.xw * nty * np * TY * 100 > (int64_t)q[2] * q[1] * 140) return -1;  // tiles must fit the image
  xg.np = np;
  xg.hh = np * TY + kh - 1;
  xg.a_bytes = HW * xg.hh * p.in.C * 2;
  xg.b_bytes = kh * kw * co * p.in.C * 2;
  xg.tiles_x = gb_make_fastdiv((uint32_t)ntx);
  xg.tiles_y = gb_make_fastdiv((uint32_t)nty);
  xg.tiles_z = gb_make_fastdiv((uint32_t)q[0]);
  const int64_t ntiles = (int64_t)ntx * nty * q[0] * p.in.N;
  if (ntiles >= (1ll << 31)) return -1;
  xg.ntiles = (int)ntiles;
  CUtensorMap ma, mb;
  if (gb_tma_activation_map_narrow(p.in, p.in.C, HW, xg.hh, &ma)) return 1;
  if (gb_tma_weight_map_narrow(p.wpacked, cc.kpad, p.npad, p.in.C, co, &mb)) return 1;
  if (p.in.C == 32) return co == 32 ? launch<32, 32>(p, ma, mb, xg, st) : launch<32, 16>(p, ma, mb, xg, st);
  return co == 32 ? launch<16, 32>(p, ma, mb, xg, st) : launch<16, 16>(p, ma, mb, xg, st);
}
