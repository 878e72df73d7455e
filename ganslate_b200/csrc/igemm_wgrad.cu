// Weight-gradient implicit GEMM on tcgen05 tensor cores.
//
//   dW[128 rows x BN cols] (TMEM, fp32) += P[64 px x 128 ch]^T  *  G[64 px x BN kflat]
//
// The reduction dimension is the pixel index, so with channels-last activations both operands are
// "MN-major": a tile row is one pixel's 64 consecutive channels (128 B), exactly the smem image the forward
// kernel builds, only the UMMA descriptors say MN-major.  P ("plain") is the tensor whose channels index the
// rows of dW (dOut for a convolution, the input for a transposed convolution); G ("gathered") is the other
// tensor read through the tap offsets, its column index kflat = tap*C + c is the forward kernel's K index.
// The pixel range is split across blockIdx.z and partial sums are combined with red.global.add.v4.f32.
#include <string.h>
#include "gb_common.cuh"
#include "gb_geometry.h"
#include "gb_tma.h"

namespace {

constexpr int BM = 128;
constexpr int BP = 64;                 // pixels per stage (MMA K = 16 -> 4 MMAs per stage)
constexpr int ATOM = BP * 128;         // one [64 px][64 ch] swizzled atom = 8 KB
constexpr int LAG = 2;

// RT = row tiles (of 128 plain channels) per CTA.  RT = 2 shares every gathered tile G between two accumulators:
// a stage is 32 KB (P) + 32 KB (G) for 1024 MMA cycles instead of 16 + 32 KB for 512 (the launch is bound by the
// ~42 B/clk/SM the L2 delivers, see igemm_pair.cu), at the price of the whole TMEM (2 x 256 columns).
template <int BN, int RT = 1>
struct WCfg {
  static constexpr int STAGES = (RT == 2) ? 3 : ((BN == 256) ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int P_BYTES = RT * 2 * ATOM;
  static constexpr int G_BYTES = (BN / 64) * ATOM;
  static constexpr int STAGE_BYTES = P_BYTES + G_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 1024;
  static constexpr int MIN_CTAS = 1;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct PixDivs {
  gb_fastdiv f[3];  // divide by the (D, H, W) extents of the plain operand's pixel grid
  // TMA variant: a 64-pixel K block is a th x tw patch; tile index -> (n, z, ty, tx)
  gb_fastdiv tiles_x, tiles_y, tiles_z;
  int tw, th, ntiles, use_tma;
};

template <int BN, bool TMA, int RT = 1>
__global__ void __launch_bounds__(256, WCfg<BN, RT>::MIN_CTAS) igemm_wgrad_kernel(const __grid_constant__ gb_wgrad_params p,
                                                                              const __grid_constant__ PixDivs divs,
                                                                              const __grid_constant__ CUtensorMap map_p,
                                                                              const __grid_constant__ CUtensorMap map_g,
                                                                              const __grid_constant__ CUtensorMap map_d,
                                                                              int blocks_per_split, int bulk_epilogue) {
  gb_pdl_enter();
  using C = WCfg<BN, RT>;
  static_assert(RT == 1 || TMA, "two row tiles per CTA only on the TMA-fed path");
  constexpr int STAGES = C::STAGES;
  constexpr int NG = BN / 64;  // gathered atoms per stage
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* tail = smem + STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 128);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int kt = blockIdx.x;           // column tile (kflat)
  const int rt = blockIdx.y * BM * RT; // first row (plain channel)
  const int64_t Mq = (int64_t)p.plain.N * p.plain.D * p.plain.H * p.plain.W;
  const int nblk = TMA ? divs.ntiles : (int)((Mq + BP - 1) / BP);
  const int b0 = blockIdx.z * blocks_per_split;
  const int b1 = min(nblk, b0 + blocks_per_split);
  const int KB = b1 - b0;
  if (KB <= 0) return;

  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = smem_u32(bars + STAGES);
  const uint32_t accum_bar = smem_u32(bars + 2 * STAGES);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, TMA ? 1 : 4);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<BN * RT>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (TMA && warp < 4) {
    // ------------------------------------------------------------------ TMA producer: one lane, (2 + NG) boxes/stage
    if (warp == 0 && gb_elect_one()) {   // (elect.sync in warp 0 only: the && short-circuits per warp)
      const int Cg = p.gathered.C;
      int g_c0[NG], g_dz[NG], g_dy[NG], g_dx[NG];
      bool g_ok[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int kflat = kt * BN + g * 64;
        const int tl = kflat / Cg;
        g_c0[g] = kflat - tl * Cg;
        g_ok[g] = tl < p.ntaps && kflat < p.kpad;
        g_dz[g] = g_ok[g] ? p.taps[tl][0] : 0;
        g_dy[g] = g_ok[g] ? p.taps[tl][1] : 0;
        g_dx[g] = g_ok[g] ? p.taps[tl][2] : 0;
        if (!g_ok[g]) g_c0[g] = Cg;  // channel coordinate past the tensor: the whole box is zero-filled
      }
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const int it = kb / STAGES;
        if (it > 0) mbar_wait(empty_bar + 8 * s, (it - 1) & 1);
        uint32_t t = (uint32_t)(b0 + kb);
        uint32_t u = gb_div(t, divs.tiles_x);
        const int x0 = (int)(t - u * divs.tiles_x.d) * divs.tw;
        t = u;
        u = gb_div(t, divs.tiles_y);
        const int y0 = (int)(t - u * divs.tiles_y.d) * divs.th;
        t = u;
        u = gb_div(t, divs.tiles_z);
        const int z0 = (int)(t - u * divs.tiles_z.d);
        const int n = (int)u;
        const uint32_t p_s = base + s * C::STAGE_BYTES;
        const uint32_t g_s = p_s + C::P_BYTES;
        const uint32_t bar = full_bar + 8 * s;
        mbar_expect_tx(bar, C::STAGE_BYTES);
#pragma unroll
        for (int a = 0; a < 2 * RT; ++a)  // channels past the tensor are zero-filled by the TMA unit
          tma_load_5d(p_s + a * ATOM, &map_p, bar, rt + a * 64, x0, y0, z0, n);
#pragma unroll
        for (int g = 0; g < NG; ++g)
          tma_load_5d(g_s + g * ATOM, &map_g, bar, g_c0[g], x0 * p.mul[2] + g_dx[g], y0 * p.mul[1] + g_dy[g],
                      z0 * p.mul[0] + g_dz[g], n);
      }
    }
    __syncwarp();
  } else if (RT == 1 && warp < 4) {
    const int j = tid & 7;
    const int r0 = tid >> 3;  // pixel rows r0 + 16*i, i < 4
    const __nv_bfloat16* pl = reinterpret_cast<const __nv_bfloat16*>(p.plain.ptr);
    const __nv_bfloat16* ga = reinterpret_cast<const __nv_bfloat16*>(p.gathered.ptr);
    // per-atom constants of the gathered operand: which tap / channel chunk this thread's column chunk is
    const int C8 = p.gathered.C >> 3;
    int g_toff[NG], g_d[NG];
    bool g_ok[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int k8 = kt * (BN / 8) + g * 8 + j;
      const int tl = k8 / C8;
      const int c8 = k8 - tl * C8;
      g_ok[g] = tl < p.ntaps && (k8 * 8) < p.kpad;
      int dz = 0, dy = 0, dx = 0;
      if (g_ok[g]) {
        dz = p.taps[tl][0];
        dy = p.taps[tl][1];
        dx = p.taps[tl][2];
      }
      g_toff[g] = (int)(dz * p.gathered.sz + dy * p.gathered.sy + dx * p.gathered.sx) + c8 * 8;
      g_d[g] = ((dz & 0xFF) << 16) | ((dy & 0xFF) << 8) | (dx & 0xFF);
    }
    // plain operand: channel of this thread's chunk in each of the two atoms
    int p_ch[2];
    bool p_ok[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      p_ch[a] = rt + (a * 8 + j) * 8;
      p_ok[a] = p_ch[a] < p.plain.C;
    }
    // all gathered atoms of this thread read the same tap when the column tile lies inside one tap
    bool uniform = true;
#pragma unroll
    for (int g = 1; g < NG; ++g) uniform = uniform && (g_d[g] == g_d[0]) && (g_ok[g] == g_ok[0]);
    const uint32_t dst0 = (uint32_t)r0 * 128u + (uint32_t)((j ^ (r0 & 7)) << 4);
    const uint32_t Mq32 = (uint32_t)Mq;
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % STAGES;
      const int it = kb / STAGES;
      if (it > 0) mbar_wait(empty_bar + 8 * s, (it - 1) & 1);
      const uint32_t p_s = base + s * C::STAGE_BYTES + dst0;
      const uint32_t g_s = p_s + C::P_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t m = (uint32_t)(b0 + kb) * BP + (uint32_t)(r0 + 16 * i);
        const bool pix_ok = m < Mq32;
        int poff = 0, goff = 0, gz = 0, gy = -100000, gx = 0;
        if (pix_ok) {
          gb_row r = gb_decode_row_fast(m, divs.f);
          poff = (int)gb_pix_offset(p.plain, r.n, r.qz, r.qy, r.qx);
          gz = r.qz * p.mul[0];
          gy = r.qy * p.mul[1];
          gx = r.qx * p.mul[2];
          goff = (int)gb_pix_offset(p.gathered, r.n, gz, gy, gx);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const bool ok = pix_ok && p_ok[a];
          cp_async16(p_s + a * ATOM + i * 2048, ok ? pl + (poff + p_ch[a]) : pl, ok);
        }
        if (uniform) {
          const int dz = (int)(signed char)(g_d[0] >> 16), dy = (int)(signed char)(g_d[0] >> 8),
                    dx = (int)(signed char)(g_d[0]);
          const bool ok = pix_ok && g_ok[0] && gb_in_bounds(p.gathered, gz + dz, gy + dy, gx + dx);
#pragma unroll
          for (int g = 0; g < NG; ++g) cp_async16(g_s + g * ATOM + i * 2048, ok ? ga + (goff + g_toff[g]) : ga, ok);
        } else {
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            const int dz = (int)(signed char)(g_d[g] >> 16), dy = (int)(signed char)(g_d[g] >> 8),
                      dx = (int)(signed char)(g_d[g]);
            const bool ok = pix_ok && g_ok[g] && gb_in_bounds(p.gathered, gz + dz, gy + dy, gx + dx);
            cp_async16(g_s + g * ATOM + i * 2048, ok ? ga + (goff + g_toff[g]) : ga, ok);
          }
        }
      }
      cp_async_commit();
      if (kb >= LAG) {
        cp_async_wait<LAG>();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar + 8 * ((kb - LAG) % STAGES));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      for (int kk = (KB > LAG ? KB - LAG : 0); kk < KB; ++kk) mbar_arrive(full_bar + 8 * (kk % STAGES));
    }
  } else if (warp == 4) {
    constexpr uint32_t idesc = make_idesc_bf16(BN, 1, 1);
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % STAGES;
      const int it = kb / STAGES;
      mbar_wait(full_bar + 8 * s, it & 1);
      tc_fence_after();
      if (gb_elect_one()) {
        const uint32_t p_s = base + s * C::STAGE_BYTES;
        const uint32_t g_s = p_s + C::P_BYTES;
        // MN-major: LBO = stride between 64-wide atoms, SBO = stride between groups of 8 pixels (k)
        const uint64_t bdesc = make_smem_desc(g_s, ATOM, 1024);
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const uint64_t adesc = make_smem_desc(p_s + r * 2 * ATOM, ATOM, 1024);
#pragma unroll
          for (int k = 0; k < BP / 16; ++k)
            umma_bf16(tmem_base + (uint32_t)(r * BN), adesc + (uint64_t)(k * 2048 / 16), bdesc + (uint64_t)(k * 2048 / 16),
                      idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(empty_bar + 8 * s);
      }
      __syncwarp();
    }
    if (lane == 0) umma_commit(accum_bar);
    __syncwarp();
  }

  mbar_wait(accum_bar, 0);
  tc_fence_after();
  if (TMA && bulk_epilogue) {
    // staged epilogue: the accumulator tile goes through 128B-swizzled [128 rows][32 floats] sub-tiles in the (idle)
    // operand ring and leaves as BN / 32 bulk reduce-adds into the fp32 workspace -- instead of 8192 16-byte
    // red.global.add per CTA whose lanes are 4 * kpad bytes apart (same move as the data kernel's fp32 epilogue:
    // 10.4 K -> 6.6 K cycles per tile, profiles/r02d_conv_timeline_b8.txt)
#pragma unroll 1
    for (int r = 0; r < RT; ++r) {
      const int lg = warp & 3;
      const int half = warp >> 2;
      const int row = lg * 32 + lane;
      const uint32_t rsw = (uint32_t)(row & 7);
      if (r > 0) {  // the previous row tile's bulk stores must have read the staging area
        if (tid == 0) tma_store_wait_read();
        __syncthreads();
      }
#pragma unroll 1
      for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(r * BN + c0), acc);
        tmem_ld_wait();
        uint8_t* dst = smem + (size_t)(c0 >> 5) * 16384 + (size_t)row * 128;
#pragma unroll
        for (int jq = 0; jq < 8; ++jq)
          *reinterpret_cast<uint4*>(dst + (((uint32_t)jq ^ rsw) << 4)) =
              make_uint4(acc[4 * jq], acc[4 * jq + 1], acc[4 * jq + 2], acc[4 * jq + 3]);
      }
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        for (int sub = 0; sub * 32 < BN; ++sub) {
          const int col = kt * BN + sub * 32;
          if (col >= p.kpad) break;
          tma_reduce_add_2d(&map_d, base + sub * 16384, col, rt + r * BM);
        }
        tma_store_commit();
      }
    }
    if (tid == 0) tma_store_wait_read();
  } else {
#pragma unroll 1
  for (int r = 0; r < RT; ++r) {
    const int lg = warp & 3;
    const int half = warp >> 2;
    const int row = rt + r * BM + lg * 32 + lane;
    const bool row_ok = row < p.rows;
    float* drow = p.dw + (int64_t)row * p.kpad + (int64_t)kt * BN;
    const int cbeg = half * (BN / 2);
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + BN / 2; c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(r * BN + c0), acc);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          if (kt * BN + c0 + g * 4 < p.kpad)
            red_add_v4(drow + c0 + g * 4, __uint_as_float(acc[g * 4 + 0]), __uint_as_float(acc[g * 4 + 1]),
                       __uint_as_float(acc[g * 4 + 2]), __uint_as_float(acc[g * 4 + 3]));
        }
      }
    }
  }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<BN * RT>(tmem_base);
}

template <int BN>
int launch(const gb_wgrad_params& p, cudaStream_t st) {
  using C = WCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(igemm_wgrad_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    GB_CUDA(cudaFuncSetAttribute(igemm_wgrad_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    if (BN == 256)
      GB_CUDA(cudaFuncSetAttribute(igemm_wgrad_kernel<256, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   WCfg<256, 2>::SMEM));
    attr_set = true;
  }
  const int64_t Mq = (int64_t)p.plain.N * p.plain.D * p.plain.H * p.plain.W;
  PixDivs divs;
  memset(&divs, 0, sizeof(divs));
  // TMA-fed producer: unit gather multiplier, both channel counts multiples of 64 (gathered) / 8 (plain)
  // (strided gathers travel as TMA boxes with traversal strides; knob 0 = 2 keeps them on the cp.async producer)
  const bool unit = p.mul[0] == 1 && p.mul[1] == 1 && p.mul[2] == 1;
  bool tma = g_gb_knobs[3] == 0 && gb_tma_available() && p.gathered.C % 64 == 0 && p.plain.C % 64 == 0 &&
             (unit || (g_gb_knobs[0] != 2 && p.mul[0] <= 4 && p.mul[1] <= 4 && p.mul[2] <= 4)) && p.plain.pad == 0 &&
             p.gathered.pad == 0;
  CUtensorMap map_p, map_g, map_d;
  memset(&map_p, 0, sizeof(map_p));
  memset(&map_g, 0, sizeof(map_g));
  memset(&map_d, 0, sizeof(map_d));
  if (tma) {
    int tw = 8;
    while (tw < p.plain.W && tw < 64) tw *= 2;
    while (tw * p.mul[2] > 256) tw /= 2;
    const int th = BP / tw;
    divs.tw = tw;
    divs.th = th;
    const int ntx = gb_cdiv(p.plain.W, tw), nty = gb_cdiv(p.plain.H, th);
    divs.tiles_x = gb_make_fastdiv((uint32_t)ntx);
    divs.tiles_y = gb_make_fastdiv((uint32_t)nty);
    divs.tiles_z = gb_make_fastdiv((uint32_t)p.plain.D);
    divs.ntiles = ntx * nty * p.plain.D * p.plain.N;
    divs.use_tma = 1;
    if (gb_tma_activation_map(p.plain, tw, th, &map_p) ||
        gb_tma_activation_map(p.gathered, tw, th, &map_g, p.mul, p.gathered_c_valid))
      return 1;
  }
  // bulk reduce-add epilogue (knob 12 = 4: per-thread red.global.add): the staged tile needs BN * 512 bytes of the ring
  const int rows_pad = gb_cdiv(p.rows, BM) * BM;
  const int bulk = tma && g_gb_knobs[12] != 4 && (size_t)BN * 512 <= (size_t)C::STAGES * C::STAGE_BYTES ? 1 : 0;
  if (bulk && gb_tma_f32_matrix_map(p.dw, p.kpad, rows_pad, BM, &map_d)) return 1;
  const int nblk = tma ? divs.ntiles : gb_cdiv(Mq, BP);
  // two row tiles per CTA: knob 12 = 2 only.  Measured slower than one (42.5 us vs 37.7 us on the residual-block
  // layer at batch 8): the launch is bound by the MMA's shared-memory operand reads, which RT = 2 does not reduce,
  // while it doubles the splits (and red.global.add traffic) needed to fill 148 SMs.
  const bool rt2 = tma && BN == 256 && p.rows > BM && g_gb_knobs[12] == 2 &&
                   (int64_t)gb_cdiv(p.kpad, BN) * gb_cdiv(p.rows, 2 * BM) * (nblk / 4) >= 96;
  const int tiles = gb_cdiv(p.kpad, BN) * gb_cdiv(p.rows, rt2 ? 2 * BM : BM);
  int splits = p.splits;
  if (splits <= 0) {
    // one wave of the 148 SMs (every split costs a full tile of red.global.add traffic in the epilogue), and at
    // least 4 pixel blocks per CTA
    splits = 148 / tiles;
    if (splits > nblk / 4) splits = nblk / 4;
    if (splits < 1) splits = 1;
  }
  const int bps = gb_cdiv(nblk, splits);
  splits = gb_cdiv(nblk, bps);
  dim3 grid(gb_cdiv(p.kpad, BN), gb_cdiv(p.rows, rt2 ? 2 * BM : BM), splits);
  divs.f[0] = gb_make_fastdiv((uint32_t)p.plain.D);
  divs.f[1] = gb_make_fastdiv((uint32_t)p.plain.H);
  divs.f[2] = gb_make_fastdiv((uint32_t)p.plain.W);
  GB_CHECK(tma || p.gathered_c_valid == 0, "gb_conv_wgrad: gathered_c_valid (pixel-window views) needs the TMA path");
  if (rt2)
    gb_klaunch(igemm_wgrad_kernel<256, true, 2>, grid, 256, WCfg<256, 2>::SMEM, st, p, divs, map_p, map_g, map_d, bps, bulk);
  else if (tma)
    gb_klaunch(igemm_wgrad_kernel<BN, true>, grid, 256, C::SMEM, st, p, divs, map_p, map_g, map_d, bps, bulk);
  else
    gb_klaunch(igemm_wgrad_kernel<BN, false>, grid, 256, C::SMEM, st, p, divs, map_p, map_g, map_d, bps, 0);
  g_gb_knobs[14] = rt2 ? 2 : (tma ? 1 : 0);  // read-back slot: which wgrad variant served the last call (tests)
  GB_LAUNCH_CHECK();
  return 0;
}

int64_t view_max_offset(const gb_view& v) {
  return (int64_t)(v.N - 1) * v.sn + (int64_t)(v.D - 1) * v.sz + (int64_t)(v.H - 1 + v.pad) * v.sy +
         (int64_t)(v.W - 1 + v.pad) * v.sx + v.C;
}

}  // namespace

int gb_conv_wgrad_narrow(const gb_wgrad_params& p, cudaStream_t st);  // igemm_wgrad_narrow.cu: -1 = not applicable

extern "C" int gb_conv_wgrad(const gb_wgrad_params* pp, void* stream) {
  const gb_wgrad_params& p = *pp;
  GB_CHECK(p.plain.ptr && p.gathered.ptr && p.dw, "gb_conv_wgrad: null pointer");
  GB_CHECK(p.plain.C % 8 == 0 && p.gathered.C % 8 == 0, "gb_conv_wgrad: channel counts must be multiples of 8");
  GB_CHECK(p.kpad % 64 == 0 && p.ntaps * p.gathered.C <= p.kpad, "gb_conv_wgrad: bad kpad %d", p.kpad);
  GB_CHECK(p.ntaps >= 1 && p.ntaps <= GB_MAX_TAPS, "gb_conv_wgrad: bad tap count %d", p.ntaps);
  GB_CHECK(p.rows >= 1 && p.rows <= p.plain.C, "gb_conv_wgrad: bad row count %d", p.rows);
  GB_CHECK(p.plain.N == p.gathered.N, "gb_conv_wgrad: batch mismatch");
  GB_CHECK((int64_t)p.plain.N * p.plain.D * p.plain.H * p.plain.W < (1ll << 31) - 64, "gb_conv_wgrad: too many pixels");
  GB_CHECK(view_max_offset(p.plain) < (1ll << 31) && view_max_offset(p.gathered) < (1ll << 31),
           "gb_conv_wgrad: tensor too large for 32-bit offsets");
  GB_CHECK(((uintptr_t)p.plain.ptr & 15) == 0 && ((uintptr_t)p.gathered.ptr & 15) == 0 && ((uintptr_t)p.dw & 15) == 0,
           "gb_conv_wgrad: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  {
    const int r = gb_conv_wgrad_narrow(p, st);  // 16 / 32 channels on both sides, dense tap box: igemm_wgrad_narrow.cu
    if (r >= 0) return r;
  }
  int bn = p.kpad >= 256 ? 256 : (p.kpad >= 128 ? 128 : 64);
  if (g_gb_knobs[2] > 0) bn = g_gb_knobs[2];
  switch (bn) {
    case 64: return launch<64>(p, st);
    case 128: return launch<128>(p, st);
    case 256: return launch<256>(p, st);
  }
  GB_CHECK(false, "gb_conv_wgrad: bad tile width %d", bn);
}
