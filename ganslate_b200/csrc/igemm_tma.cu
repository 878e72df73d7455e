// TMA-fed variant of the implicit-GEMM convolution (forward / data gradient / transposed classes) for the
// dominant case: unit gather multiplier (stride-1 convolutions and every parity class of a transposed or
// strided-dgrad convolution) and a channel count that is a multiple of 64.
//
// A GEMM row tile is a TH x TW patch of output pixels (TH*TW <= 128, any TW: an 11 x 11 patch covers the 66 x 66
// reflection-padded domain of the residual-block data gradients with 36 tiles per image instead of the 45 that
// 16 x 8 patches need) of one image / depth slice.  For tap t and
// 64-channel chunk c the A operand is then one 5-D TMA box {64 ch, TW, TH, 1, 1} of the channels-last input at
// pixel offset d_t -- the hardware does the address generation, zero-fills outside the image (= zero padding)
// and writes the 128B-swizzled K-major tile the UMMA descriptor expects.  B (packed weights) is a 2-D box
// {64, BN}.  One elected thread issues both loads per stage; there is no per-thread gather work at all, which
// is what bounded the cp.async kernel (igemm_data.cu: ~12 % tensor-pipe, producers issue-latency bound).
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), all 8 warps = epilogue.
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include "gb_common.cuh"
#include "gb_geometry.h"
#include "gb_epilogue.cuh"
#include "gb_tma.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;

template <int BN>
struct TCfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // TMA latency is ~1 us: what matters is bytes in flight per SM, so use every stage that fits in ~200 KB
  // (one CTA per SM)
  static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 8 ? 8 : (200 * 1024 / STAGE_BYTES);
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 1024;
  static constexpr int MIN_CTAS = 1;
};

struct TileGeom {
  gb_fastdiv tiles_x, tiles_y, tiles_z;  // decode tile index -> (n, z, ty, tx)
  int tw, th;
  gb_fastdiv div_tw;                     // tile row -> (h, w) = (row / tw, row % tw); rows >= tw * th are unused
  int ntx, nty;
  int ntiles;                            // per class (max over classes is the grid)
  int nstages;                           // depth of the smem ring (<= TCfg::STAGES); small rings let two CTAs share an SM
  int store_mode;                        // 0 = per-thread global stores (gb_conv_epilogue); TMA-store epilogue: 1 = bf16,
                                         // 2 = fp32, 3 = fp32 accumulated into the destination (bulk reduce-add)
  // bring-up instrumentation (tools/conv_timeline.py; both 0 in every ordinary launch):
  int mode;                              // knob 30 (bits): 1 = producer skips the loads, 2 = issuer skips the MMAs, 4 = no epilogue
  unsigned long long* ts;                // gb_debug_timeline(): 16 words per CTA (smid, globaltimer, clock64 stamps)
};

__device__ __forceinline__ void ts_put(const TileGeom& tg, int slot) {
  if (tg.ts != nullptr) {
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    tg.ts[16ull * cta + slot] = (unsigned long long)clock64();
  }
}


// ------------------------------------------------------------------------------------------------- staged epilogue
// TMEM accumulator -> registers (+bias, activation) -> 128B-swizzled staging tiles in the (now idle) operand ring ->
// global memory as whole 128-byte rows, either by the warps (lane = 16-byte chunk of a row, four rows per warp
// instruction: store modes 4 / 5 / 6) or by ONE bulk tensor store per 128-byte channel group (modes 1 / 2 / 3).
// The per-thread version (gb_conv_epilogue, mode 0) issues 16-byte stores 512 B apart from 110-register threads and
// measured 6.2 K (bf16) / 10.4 K (fp32) cycles per 128 x 256 tile against an 18.4 K cycle main loop
// (profiles/r02b_conv_timeline_b8.txt).  InstanceNorm statistics are summed from the staged bf16 tile: 16-byte reads,
// eight channels per thread, shared-memory atomics, one global atomic per channel and CTA.
//   staging layout: sub-tile s (64 bf16 / 32 fp32 channels) at s * 16 KB; row r (pixel h * tw + w) at r * 128 B; its
//   16-byte chunk j at ((j ^ (r & 7)) << 4) -- what CU_TENSOR_MAP_SWIZZLE_128B expects, and conflict-free both for a
//   warp whose lanes are consecutive rows (staging) and for one whose lanes are the chunks of four rows (write-out).
struct EpiCoord {
  int x0, y0, n;          // first q-grid pixel of the tile, image
  int oz;                 // output z coordinate
  int qh, qw;             // q-grid extents (rows / columns that exist)
  int cx, cy;             // output coordinates of the tile's first pixel (TMA store)
};

template <int BN, int ACT, bool FP32>
__device__ __forceinline__ void stage_tile(const gb_conv_params& p, uint32_t tmem_base, int lg, int half, int lane,
                                           bool have_acc, bool row_ok, const float* bias_s, uint8_t* stage) {
  const int row = lg * 32 + lane;
  const uint32_t rsw = (uint32_t)(row & 7);
  constexpr int CH = 32;
  constexpr int COLS_PER_HALF = BN / 2;
#pragma unroll 1
  for (int c0 = half * COLS_PER_HALF; c0 < (half + 1) * COLS_PER_HALF; c0 += CH) {
    uint32_t acc[CH];
    if (have_acc) {
      tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, acc);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i) acc[i] = 0u;
    }
    float v[CH];
    const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);
#pragma unroll
    for (int jq = 0; jq < CH / 4; ++jq) {
      const float4 b = b4[jq];
      v[4 * jq + 0] = __uint_as_float(acc[4 * jq + 0]) + b.x;
      v[4 * jq + 1] = __uint_as_float(acc[4 * jq + 1]) + b.y;
      v[4 * jq + 2] = __uint_as_float(acc[4 * jq + 2]) + b.z;
      v[4 * jq + 3] = __uint_as_float(acc[4 * jq + 3]) + b.w;
    }
    if constexpr (ACT != GB_ACT_NONE) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if constexpr (ACT == GB_ACT_TANH) v[i] = tanhf(v[i]);
        else if constexpr (ACT == GB_ACT_LEAKY) v[i] = v[i] > 0.f ? v[i] : v[i] * p.act_slope;
        else v[i] = fmaxf(v[i], 0.f);
      }
    }
    if constexpr (FP32) {
      // 32 fp32 columns = one 128-byte row of sub-tile c0 / 32
      uint8_t* dst = stage + (size_t)(c0 >> 5) * 16384 + (size_t)row * 128;
#pragma unroll
      for (int jq = 0; jq < 8; ++jq)
        *reinterpret_cast<float4*>(dst + (((uint32_t)jq ^ rsw) << 4)) =
            make_float4(v[4 * jq], v[4 * jq + 1], v[4 * jq + 2], v[4 * jq + 3]);
    } else {
      if (!row_ok) {   // rows outside the image: zeros for the statistics (never written to global memory)
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] = 0.f;
      }
      // 32 bf16 columns = half a 128-byte row (chunks jb .. jb + 3) of sub-tile c0 / 64
      uint8_t* dst = stage + (size_t)(c0 >> 6) * 16384 + (size_t)row * 128;
      const uint32_t jb = (uint32_t)((c0 & 63) >> 3);
#pragma unroll
      for (int jq = 0; jq < 4; ++jq) {
        uint4 o;
        o.x = pack_bf16x2(v[8 * jq + 0], v[8 * jq + 1]);
        o.y = pack_bf16x2(v[8 * jq + 2], v[8 * jq + 3]);
        o.z = pack_bf16x2(v[8 * jq + 4], v[8 * jq + 5]);
        o.w = pack_bf16x2(v[8 * jq + 6], v[8 * jq + 7]);
        *reinterpret_cast<uint4*>(dst + (((jb + (uint32_t)jq) ^ rsw) << 4)) = o;
      }
    }
  }
}

template <int BN>
__device__ __forceinline__ void staged_epilogue(const gb_conv_params& p, const TileGeom& tg, const CUtensorMap* map_o,
                                                uint32_t tmem_base, int warp, int lane, bool have_acc, bool row_ok,
                                                int n0, const float* bias_s, uint8_t* stage, float* sacc,
                                                const EpiCoord& ec, const gb_conv_class& cc) {
  const int tid = warp * 32 + lane;
  const int lg = warp & 3, half = warp >> 2;
  const int mode = tg.store_mode;
  const bool fp32 = (mode == 2 || mode == 3 || mode == 5 || mode == 6);
  const bool use_tma = mode <= 3;
  const bool want_stats = !fp32 && p.stats != nullptr;
  static_assert(BN >= 64, "staged epilogue needs BN >= 64");
  (void)sacc;
  if (fp32) {
    stage_tile<BN, GB_ACT_NONE, true>(p, tmem_base, lg, half, lane, have_acc, row_ok, bias_s, stage);
  } else {
    switch (p.act) {
      case GB_ACT_TANH: stage_tile<BN, GB_ACT_TANH, false>(p, tmem_base, lg, half, lane, have_acc, row_ok, bias_s, stage); break;
      case GB_ACT_LEAKY: stage_tile<BN, GB_ACT_LEAKY, false>(p, tmem_base, lg, half, lane, have_acc, row_ok, bias_s, stage); break;
      case GB_ACT_RELU: stage_tile<BN, GB_ACT_RELU, false>(p, tmem_base, lg, half, lane, have_acc, row_ok, bias_s, stage); break;
      default: stage_tile<BN, GB_ACT_NONE, false>(p, tmem_base, lg, half, lane, have_acc, row_ok, bias_s, stage); break;
    }
  }
  if (tid == 64) ts_put(tg, 8);    // tile staged
  if (use_tma) fence_proxy_async();   // staging writes (generic proxy) -> visible to the bulk store (async proxy)
  tc_fence_before();
  __syncthreads();
  if (tid == 64) ts_put(tg, 9);
  const int rows = tg.tw * tg.th;
  if (use_tma) {
    if (tid == 0) {
      const int inner = fp32 ? 32 : 64;
      const uint32_t s0 = smem_u32(stage);
      for (int sub = 0; sub * inner < BN; ++sub) {
        const int c = n0 + sub * inner;
        if (c >= p.out.C) break;
        if (mode == 3) tma_reduce_add_5d(map_o, s0 + sub * 16384, c, ec.cx, ec.cy, ec.oz, ec.n);
        else tma_store_5d(map_o, s0 + sub * 16384, c, ec.cx, ec.cy, ec.oz, ec.n);
      }
      tma_store_commit();
    }
  } else {
    // write-out by the warps: lane = (row of a quad, 16-byte chunk); a warp instruction moves four whole 128-byte rows
    const int lr = lane >> 3;
    const uint32_t jc = (uint32_t)(lane & 7);
    const int esz = fp32 ? 4 : 2;
    const int epc = 16 / esz;                  // elements per 16-byte chunk
    const int inner = 128 / esz;               // channels per sub-tile
    uint8_t* obase = reinterpret_cast<uint8_t*>(p.out.ptr);
    for (int r = warp * 4 + lr; r < rows; r += 32) {
      const int h = (int)gb_div((uint32_t)r, tg.div_tw), w = r - h * tg.tw;
      const int qy = ec.y0 + h, qx = ec.x0 + w;
      if (qy >= ec.qh || qx >= ec.qw) continue;
      const int64_t off = gb_pix_offset(p.out, ec.n, ec.oz, qy * p.out_mul[1] + cc.off[1], qx * p.out_mul[2] + cc.off[2]);
      const uint8_t* srow = stage + (size_t)r * 128 + ((jc ^ (uint32_t)(r & 7)) << 4);
#pragma unroll 1
      for (int sub = 0; sub * inner < BN; ++sub) {
        const int col = n0 + sub * inner + (int)jc * epc;
        if (col >= p.out.C) break;
        const uint4 q4 = *reinterpret_cast<const uint4*>(srow + (size_t)sub * 16384);
        uint8_t* dst = obase + (off + col) * esz;
        if (mode == 6) {
          float4 o = *reinterpret_cast<float4*>(dst);
          o.x += __uint_as_float(q4.x); o.y += __uint_as_float(q4.y);
          o.z += __uint_as_float(q4.z); o.w += __uint_as_float(q4.w);
          *reinterpret_cast<float4*>(dst) = o;
        } else {
          *reinterpret_cast<uint4*>(dst) = q4;
        }
      }
    }
  }
  if (tid == 64) ts_put(tg, 10);   // write-out issued
  if (want_stats) {
    // thread -> (16-byte chunk cg of a row = 8 channels, row group rg); rows rg, rg + RG, ... of the tile
    constexpr int CPR = BN / 8;        // chunks per row over all sub-tiles
    constexpr int RG = 256 / CPR;      // row groups
    const int cg = tid % CPR, rg = tid / CPR;
    const uint8_t* src = stage + (size_t)(cg >> 3) * 16384;
    const uint32_t jc = (uint32_t)(cg & 7);
    float s1[8], s2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
    for (int r = rg; r < rows; r += RG) {
      const uint4 q4 = *reinterpret_cast<const uint4*>(src + (size_t)r * 128 + ((jc ^ (uint32_t)(r & 7)) << 4));
      float2 f;
      f = unpack_bf16x2(q4.x); s1[0] += f.x; s2[0] = fmaf(f.x, f.x, s2[0]); s1[1] += f.y; s2[1] = fmaf(f.y, f.y, s2[1]);
      f = unpack_bf16x2(q4.y); s1[2] += f.x; s2[2] = fmaf(f.x, f.x, s2[2]); s1[3] += f.y; s2[3] = fmaf(f.y, f.y, s2[3]);
      f = unpack_bf16x2(q4.z); s1[4] += f.x; s2[4] = fmaf(f.x, f.x, s2[4]); s1[5] += f.y; s2[5] = fmaf(f.y, f.y, s2[5]);
      f = unpack_bf16x2(q4.w); s1[6] += f.x; s2[6] = fmaf(f.x, f.x, s2[6]); s1[7] += f.y; s2[7] = fmaf(f.y, f.y, s2[7]);
    }
    // partial sums of the RG row groups -> scratch behind the staged tile (plain stores: every (rg, channel) has one
    // owner), then one thread per channel adds the RG partials and issues the CTA's single atomic per moment
    float* part = reinterpret_cast<float*>(stage + (size_t)(BN / 64) * 16384);   // [RG][BN][2]
#pragma unroll
    for (int e = 0; e < 8; e += 2)
      *reinterpret_cast<float4*>(part + ((size_t)rg * BN + cg * 8 + e) * 2) = make_float4(s1[e], s2[e], s1[e + 1], s2[e + 1]);
    __syncthreads();
    for (int i = tid; i < BN; i += 256) {
      const int col = n0 + i;
      if (col < p.ncols) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < RG; ++k) {
          const float2 v = *reinterpret_cast<const float2*>(part + ((size_t)k * BN + i) * 2);
          a += v.x;
          b += v.y;
        }
        float* dst = p.stats + ((int64_t)ec.n * p.out.C + col) * 2;
        atomicAdd(dst, a);
        atomicAdd(dst + 1, b);
      }
    }
  }
  if (tid == 64) ts_put(tg, 11);   // statistics done
  if (use_tma && tid == 0) tma_store_wait_read();
}

template <int BN>
__global__ void __launch_bounds__(256, TCfg<BN>::MIN_CTAS)
igemm_tma_kernel(const __grid_constant__ gb_conv_params p, const __grid_constant__ CUtensorMap map_a,
                 const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_o,
                 const __grid_constant__ TileGeom tg) {
  gb_pdl_enter();
  using C = TCfg<BN>;
  constexpr int MAXS = 8;  // barrier slots (TCfg::STAGES <= 8)
  const int STAGES = tg.nstages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* tail = smem + STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);  // full[STAGES], empty[STAGES], accum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 144);
  int8_t* taps_s = reinterpret_cast<int8_t*>(tail + 192);
  __shared__ __align__(16) float bias_s[BN];
  __shared__ float stat_s[BN >= 64 ? 2 * BN : 2];   // TMA-store epilogue: per-channel (sum, sum^2) of this tile
  // Narrow tiles (BN < 64): the InstanceNorm statistics of the four row warps meet here and leave the CTA as one atomic
  // per column and moment.  Measured (tools/tiny_k_probe.py): the 64 -> 16 transposed convolution of the V-Net on
  // 32 x 256 x 256 voxels takes 346 us with per-warp global atomics (2 M of them on 32 addresses), 125 us without any.
  __shared__ float sst_narrow[BN < 64 ? 2 * BN : 2];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cls = blockIdx.z;
  const gb_conv_class& cc = p.cls[cls];
  int q[3];
  gb_class_extents(p, cls, q);
  // tile -> (n, z, ty, tx); tiles of classes with smaller extents simply fall outside and exit
  uint32_t t = blockIdx.x;
  uint32_t u = gb_div(t, tg.tiles_x);
  const int tx = (int)(t - u * tg.tiles_x.d);
  t = u;
  u = gb_div(t, tg.tiles_y);
  const int ty = (int)(t - u * tg.tiles_y.d);
  t = u;
  u = gb_div(t, tg.tiles_z);
  const int z0 = (int)(t - u * tg.tiles_z.d);
  const int n = (int)u;
  const int x0 = tx * tg.tw, y0 = ty * tg.th;
  if (n >= p.in.N || z0 >= q[0] || y0 >= q[1] || x0 >= q[2]) return;
  const int n0 = blockIdx.y * BN;
  const int chunks = p.in.C >> 6;
  const int KB = cc.ntaps * chunks;
  if (tg.ts != nullptr && tid == 0) {
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    unsigned smid;
    unsigned long long gt;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tg.ts[16ull * cta + 0] = smid;
    tg.ts[16ull * cta + 1] = gt;
    tg.ts[16ull * cta + 2] = (unsigned long long)clock64();
  }

  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = smem_u32(bars + MAXS);
  const uint32_t accum_bar = smem_u32(bars + 2 * MAXS);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(smem_u32(tmem_slot));
  for (int i = tid; i < cc.ntaps; i += 256)
    *reinterpret_cast<uint32_t*>(taps_s + 4 * i) = *reinterpret_cast<const uint32_t*>(p.taps[cc.tap_begin + i]);
  for (int i = tid; i < BN; i += 256) bias_s[i] = (p.bias != nullptr && n0 + i < p.ncols) ? p.bias[n0 + i] : 0.f;
  if (BN < 64)
    for (int i = tid; i < 2 * BN; i += 256) sst_narrow[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 64) ts_put(tg, 3);   // setup done

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one lane; elect.sync keeps the
    // descriptor / coordinate arithmetic of the region in the uniform datapath, see gb_elect_one)
    if (gb_elect_one()) {
      // packed weights of this class start at row (w_offset / kpad) of the 2-D weight map built per class
      int s = 0, it = 0;
      for (int tl = 0; tl < cc.ntaps; ++tl) {
        const int dz = taps_s[4 * tl + 0], dy = taps_s[4 * tl + 1], dx = taps_s[4 * tl + 2];
        for (int c = 0; c < chunks; ++c) {
          if (it > 0) mbar_wait(empty_bar + 8 * s, (it - 1) & 1);
          const uint32_t a_s = base + s * C::STAGE_BYTES;
          const uint32_t b_s = a_s + A_BYTES;
          const uint32_t bar = full_bar + 8 * s;
          if (tg.mode & 1) {
            mbar_arrive(bar);
          } else {
            mbar_expect_tx(bar, (uint32_t)(tg.tw * tg.th * 128 + C::B_BYTES));
            tma_load_5d(a_s, &map_a, bar, c * 64, x0 * p.in_mul[2] + dx, y0 * p.in_mul[1] + dy, z0 * p.in_mul[0] + dz, n);
            tma_load_2d(b_s, &map_b, bar, tl * p.in.C + c * 64, cls * p.npad + n0);
          }
          if (++s == STAGES) {
            s = 0;
            ++it;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BN, 0, 0);
    int s = 0, it = 0;
    for (int kb = 0; kb < KB; ++kb) {
      mbar_wait(full_bar + 8 * s, it & 1);
      tc_fence_after();
      if (kb == 0 && lane == 0) ts_put(tg, 4);   // first operands have landed
      if (tg.mode == 2) {
        if (lane == 0) mbar_arrive(empty_bar + 8 * s);
      } else if (gb_elect_one()) {
        const uint32_t a_s = base + s * C::STAGE_BYTES;
        const uint32_t b_s = a_s + A_BYTES;
        const uint64_t adesc = make_smem_desc(a_s, 16, 1024);
        const uint64_t bdesc = make_smem_desc(b_s, 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
        umma_commit(empty_bar + 8 * s);
      }
      __syncwarp();
      if (++s == STAGES) {
        s = 0;
        ++it;
      }
    }
    if (lane == 0) ts_put(tg, 5);                // every MMA issued
    if (lane == 0 && KB > 0) {
      if (tg.mode == 2) mbar_arrive(accum_bar); else umma_commit(accum_bar);
    }
    __syncwarp();
  }

  // -------------------------------------------------------------------- epilogue (all warps)
  if (KB > 0) {
    mbar_wait(accum_bar, 0);
    tc_fence_after();
  }
  if (tid == 64) ts_put(tg, 6);                  // accumulator complete
  if (!(tg.mode & 4)) {
    const int row = (warp & 3) * 32 + lane;
    const int h = (int)gb_div((uint32_t)row, tg.div_tw), w = row - h * tg.tw;
    const int qy = y0 + h, qx = x0 + w;
    const bool row_ok = h < tg.th && qy < q[1] && qx < q[2];
    bool done = false;
    if constexpr (BN >= 64) {
      if (tg.store_mode != 0) {
        // (the operand ring is idle once the accumulator is complete: staging tiles of the write-out)
        EpiCoord ec;
        ec.x0 = x0, ec.y0 = y0, ec.n = n, ec.oz = z0 * p.out_mul[0] + cc.off[0];
        ec.qh = q[1], ec.qw = q[2];
        ec.cx = x0 * p.out_mul[2] + cc.off[2], ec.cy = y0 * p.out_mul[1] + cc.off[1];
        staged_epilogue<BN>(p, tg, &map_o, tmem_base, warp, lane, KB > 0, row_ok, n0, bias_s, smem, stat_s, ec, cc);
        done = true;
      }
    }
    if (!done) {
      int64_t ooff = 0;
      if (row_ok)
        ooff = gb_pix_offset(p.out, n, z0 * p.out_mul[0] + cc.off[0], qy * p.out_mul[1] + cc.off[1],
                             qx * p.out_mul[2] + cc.off[2]);
      // (stage 0 of the ring is free once the accumulator is complete: scratch of the CTA-level statistics sum)
      gb_conv_epilogue<BN>(p, tmem_base, warp, lane, KB > 0, row_ok, ooff, n0, bias_s, n, reinterpret_cast<float*>(smem),
                           (BN < 64 && p.stats != nullptr && !p.out_fp32) ? sst_narrow : nullptr);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (BN < 64 && p.stats != nullptr && !p.out_fp32 && !(tg.mode & 4)) {
    for (int i = tid; i < 2 * BN; i += 256)
      if (n0 + (i >> 1) < p.ncols) atomicAdd(p.stats + ((int64_t)n * p.out.C + n0 + (i >> 1)) * 2 + (i & 1), sst_narrow[i]);
  }
  if (tid == 64) ts_put(tg, 7);                  // epilogue done
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// ============================================================================================ persistent kernel
// One CTA per SM walks a CONTIGUOUS range of (class, column block, pixel tile) items.  The per-CTA fixed costs of the
// kernel above (barrier init, TMEM allocation, first-load latency: ~2.5 K cycles) are paid once, and the epilogue of
// item i -- bounded by the 64 B/clk TMEM read and the ~32 B/clk an SM can write to L2: 4 - 7 K cycles for a 128 x 256
// tile, as long as the 18.4 K cycle main loop of a residual-block tile and longer than the whole main loop of the
// short-K layers (profiles/r02d_conv_timeline_b8.txt) -- runs under the main loop of item i + 1:
//   warp 0      TMA producer, operand ring continuous across items
//   warp 1      MMA issuer; two TMEM accumulators (acc_full / acc_empty barriers)
//   warps 2-9   epilogue: TMEM -> registers (+bias, activation, bf16) -> 16 KB staging tile -> coalesced 128-byte rows.
//               Two warps share each TMEM lane group (32 rows) and split the columns; a "round" stages 128 bytes of
//               channels per row, the pair meets at a 64-thread named barrier and writes its 32 rows out, four rows
//               per warp instruction.  InstanceNorm statistics: butterfly column sums of the bf16-rounded values kept
//               in REGISTERS across the items of one image (lane = column), flushed with one atomic per column when
//               the image / column block changes -- consecutive items of a CTA are neighbouring tiles of one image.
template <int BN>
struct PCfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING = 16384;
  static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 8 ? 8 : (200 * 1024 / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM = STAGES * STAGE_BYTES + STAGING + 2048;
};

struct PersGeom {
  TileGeom tg;
  int ncb;                    // column blocks of BN output channels
  int total;                  // items = nclass * ncb * tg.ntiles
  gb_fastdiv div_tiles, div_ncb;
};

struct PItem {
  int cls, n0, n, z0, y0, x0, KB, qh, qw;
};

__device__ __forceinline__ bool pers_decode(const gb_conv_params& p, const PersGeom& pg, int item, int bn, PItem& it) {
  uint32_t t = (uint32_t)item;
  uint32_t u = gb_div(t, pg.div_tiles);
  uint32_t tile = t - u * pg.div_tiles.d;
  uint32_t c = gb_div(u, pg.div_ncb);
  it.n0 = (int)(u - c * pg.div_ncb.d) * bn;
  it.cls = (int)c;
  int q[3];
  gb_class_extents(p, it.cls, q);
  t = tile;
  u = gb_div(t, pg.tg.tiles_x);
  it.x0 = (int)(t - u * pg.tg.tiles_x.d) * pg.tg.tw;
  t = u;
  u = gb_div(t, pg.tg.tiles_y);
  it.y0 = (int)(t - u * pg.tg.tiles_y.d) * pg.tg.th;
  t = u;
  u = gb_div(t, pg.tg.tiles_z);
  it.z0 = (int)(t - u * pg.tg.tiles_z.d);
  it.n = (int)u;
  it.qh = q[1];
  it.qw = q[2];
  it.KB = p.cls[it.cls].ntaps * (p.in.C >> 6);
  return it.n < p.in.N && it.z0 < q[0] && it.y0 < q[1] && it.x0 < q[2];
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one item's epilogue for one warp.  FP32: rounds of 32 fp32 columns (16 per warp), else rounds of 64 bf16 columns (32
// per warp).  st1 / st2: running column sums of this warp's columns (bf16 path with statistics).
template <int BN, int ACT, bool FP32, bool ACCUM>
__device__ __forceinline__ void pers_epilogue_item(const gb_conv_params& p, const PersGeom& pg, const PItem& it,
                                                   uint32_t tmem_acc, uint32_t acc_empty_bar, int lg, int half, int lane,
                                                   uint8_t* stage, bool want_stats, float (&st1)[BN / 64],
                                                   float (&st2)[BN / 64]) {
  constexpr int ROUNDS = FP32 ? BN / 32 : BN / 64;
  constexpr int ESZ = FP32 ? 4 : 2;
  const gb_conv_class& cc = p.cls[it.cls];
  const int row = lg * 32 + lane;
  const uint32_t rsw = (uint32_t)(row & 7);
  const int h = (int)gb_div((uint32_t)row, pg.tg.div_tw), w = row - h * pg.tg.tw;
  const bool row_ok = h < pg.tg.th && it.y0 + h < it.qh && it.x0 + w < it.qw;
  uint8_t* srow = stage + (size_t)row * 128;
  // write-out role of this thread: chunk j of rows (t2 >> 3) + 8k of the pair's 32 rows
  const int t2 = half * 32 + lane;
  const uint32_t jc = (uint32_t)(t2 & 7);
  int64_t woff[4];
  bool wok[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r2 = lg * 32 + (t2 >> 3) + 8 * k;
    const int h2 = (int)gb_div((uint32_t)r2, pg.tg.div_tw), w2 = r2 - h2 * pg.tg.tw;
    const int qy = it.y0 + h2, qx = it.x0 + w2;
    wok[k] = h2 < pg.tg.th && qy < it.qh && qx < it.qw;
    woff[k] = wok[k] ? gb_pix_offset(p.out, it.n, it.z0 * p.out_mul[0] + cc.off[0], qy * p.out_mul[1] + cc.off[1],
                                     qx * p.out_mul[2] + cc.off[2])
                     : 0;
  }
  uint8_t* obase = reinterpret_cast<uint8_t*>(p.out.ptr);
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    if constexpr (FP32) {
      const int c0 = r * 32 + half * 16;
      uint32_t acc[16];
      tmem_ld16(tmem_acc + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, acc);
      tmem_ld_wait();
      if (r == ROUNDS - 1) {   // this warp has read its part of the accumulator: hand it back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty_bar);
      }
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = __uint_as_float(acc[i]);
        if (p.bias != nullptr && it.n0 + c0 + i < p.ncols) v[i] += __ldg(p.bias + it.n0 + c0 + i);
      }
#pragma unroll
      for (int jq = 0; jq < 4; ++jq)
        *reinterpret_cast<float4*>(srow + (((uint32_t)(4 * half + jq)) ^ rsw) * 16) =
            make_float4(v[4 * jq], v[4 * jq + 1], v[4 * jq + 2], v[4 * jq + 3]);
    } else {
      const int c0 = r * 64 + half * 32;
      uint32_t acc[32];
      tmem_ld32(tmem_acc + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, acc);
      tmem_ld_wait();
      if (r == ROUNDS - 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty_bar);
      }
      float v[32];
      if (p.bias != nullptr) {
        if (it.n0 + c0 + 32 <= p.ncols) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + it.n0 + c0);
#pragma unroll
          for (int jq = 0; jq < 8; ++jq) {
            const float4 b = __ldg(b4 + jq);
            v[4 * jq + 0] = __uint_as_float(acc[4 * jq + 0]) + b.x;
            v[4 * jq + 1] = __uint_as_float(acc[4 * jq + 1]) + b.y;
            v[4 * jq + 2] = __uint_as_float(acc[4 * jq + 2]) + b.z;
            v[4 * jq + 3] = __uint_as_float(acc[4 * jq + 3]) + b.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[i] = __uint_as_float(acc[i]) + (it.n0 + c0 + i < p.ncols ? __ldg(p.bias + it.n0 + c0 + i) : 0.f);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
      }
      if constexpr (ACT != GB_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if constexpr (ACT == GB_ACT_TANH) v[i] = tanhf(v[i]);
          else if constexpr (ACT == GB_ACT_LEAKY) v[i] = v[i] > 0.f ? v[i] : v[i] * p.act_slope;
          else v[i] = fmaxf(v[i], 0.f);
        }
      }
      uint32_t o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
      if (want_stats) {
        float sv[32], sq[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 f = unpack_bf16x2(o[i]);
          sv[2 * i] = f.x;
          sv[2 * i + 1] = f.y;
        }
        if (!row_ok) {
#pragma unroll
          for (int i = 0; i < 32; ++i) sv[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) sq[i] = sv[i] * sv[i];
        st1[r] += gb_warp_colsum<32>(sv, lane);
        st2[r] += gb_warp_colsum<32>(sq, lane);
      }
#pragma unroll
      for (int jq = 0; jq < 4; ++jq)
        *reinterpret_cast<uint4*>(srow + (((uint32_t)(4 * half + jq)) ^ rsw) * 16) =
            make_uint4(o[4 * jq], o[4 * jq + 1], o[4 * jq + 2], o[4 * jq + 3]);
    }
    named_bar_sync(1 + lg, 64);   // the pair's 32 rows x 128 bytes are staged
    const int col = it.n0 + r * (128 / ESZ) + (int)jc * (16 / ESZ);
    if (col < p.out.C) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (wok[k]) {
          const int r2 = lg * 32 + (t2 >> 3) + 8 * k;
          const uint4 q4 = *reinterpret_cast<const uint4*>(stage + (size_t)r2 * 128 + ((jc ^ (uint32_t)(r2 & 7)) << 4));
          uint8_t* dst = obase + (woff[k] + col) * ESZ;
          if constexpr (ACCUM) {
            float4 a = *reinterpret_cast<float4*>(dst);
            a.x += __uint_as_float(q4.x); a.y += __uint_as_float(q4.y);
            a.z += __uint_as_float(q4.z); a.w += __uint_as_float(q4.w);
            *reinterpret_cast<float4*>(dst) = a;
          } else {
            *reinterpret_cast<uint4*>(dst) = q4;
          }
        }
      }
    }
    named_bar_sync(1 + lg, 64);   // the rows have been read: the next round may overwrite them
  }
}

template <int BN>
__global__ void __launch_bounds__(320, 1)
igemm_pers_kernel(const __grid_constant__ gb_conv_params p, const __grid_constant__ CUtensorMap map_a,
                  const __grid_constant__ CUtensorMap map_b, const __grid_constant__ PersGeom pg) {
  gb_pdl_enter();
  using C = PCfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint8_t* stage = smem + STAGES * C::STAGE_BYTES;
  uint8_t* tail = stage + C::STAGING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);  // full[STAGES], empty[STAGES], acc_full[2], acc_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 8 * (2 * STAGES + 4));
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int it_begin = (int)((int64_t)blockIdx.x * pg.total / gridDim.x);
  const int it_end = (int)((int64_t)(blockIdx.x + 1) * pg.total / gridDim.x);
  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = smem_u32(bars + STAGES);
  const uint32_t accf_bar = smem_u32(bars + 2 * STAGES);
  const uint32_t acce_bar = smem_u32(bars + 2 * STAGES + 2);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(accf_bar, 1);
    mbar_init(accf_bar + 8, 1);
    mbar_init(acce_bar, 8);
    mbar_init(acce_bar + 8, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int chunks = p.in.C >> 6;
  unsigned long long* tsb = pg.tg.ts != nullptr ? pg.tg.ts + 16ull * blockIdx.x : nullptr;   // gb_debug_timeline
  if (tsb != nullptr && tid == 0) {
    unsigned smid;
    unsigned long long gt;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tsb[0] = smid;
    tsb[1] = gt;
    tsb[3] = (unsigned long long)clock64();
    tsb[2] = tsb[3];
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one lane)
    if (lane == 0) {
      uint32_t g = 0;
      for (int item = it_begin; item < it_end; ++item) {
        PItem it;
        if (!pers_decode(p, pg, item, BN, it)) continue;
        const gb_conv_class& cc = p.cls[it.cls];
        for (int tl = 0; tl < cc.ntaps; ++tl) {
          const int dz = p.taps[cc.tap_begin + tl][0], dy = p.taps[cc.tap_begin + tl][1], dx = p.taps[cc.tap_begin + tl][2];
          for (int c = 0; c < chunks; ++c, ++g) {
            const uint32_t s = g % STAGES, itn = g / STAGES;
            if (itn > 0) mbar_wait(empty_bar + 8 * s, (itn - 1) & 1);
            const uint32_t a_s = base + s * C::STAGE_BYTES;
            const uint32_t bar = full_bar + 8 * s;
            mbar_expect_tx(bar, (uint32_t)(pg.tg.tw * pg.tg.th * 128 + C::B_BYTES));
            tma_load_5d(a_s, &map_a, bar, c * 64, it.x0 * p.in_mul[2] + dx, it.y0 * p.in_mul[1] + dy,
                        it.z0 * p.in_mul[0] + dz, it.n);
            tma_load_2d(a_s + A_BYTES, &map_b, bar, tl * p.in.C + c * 64, it.cls * p.npad + it.n0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BN, 0, 0);
    uint32_t g = 0;
    int li = 0;
    for (int item = it_begin; item < it_end; ++item) {
      PItem it;
      if (!pers_decode(p, pg, item, BN, it)) continue;
      const int bsel = li & 1, use = li >> 1;
      if (use > 0) {
        mbar_wait(acce_bar + 8 * bsel, (use - 1) & 1);
        tc_fence_after();
      }
      const uint32_t tacc = tmem_base + (uint32_t)(bsel * BN);
      for (int kb = 0; kb < it.KB; ++kb, ++g) {
        const uint32_t s = g % STAGES, itn = g / STAGES;
        mbar_wait(full_bar + 8 * s, itn & 1);
        tc_fence_after();
        if (tsb != nullptr && kb == 0 && lane == 0 && li < 2) tsb[4 + 2 * li] = (unsigned long long)clock64();
        if (lane == 0) {
          const uint32_t a_s = base + s * C::STAGE_BYTES;
          const uint64_t adesc = make_smem_desc(a_s, 16, 1024);
          const uint64_t bdesc = make_smem_desc(a_s + A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(empty_bar + 8 * s);
        }
        __syncwarp();
      }
      if (lane == 0) umma_commit(accf_bar + 8 * bsel);
      if (tsb != nullptr && lane == 0 && li < 2) tsb[5 + 2 * li] = (unsigned long long)clock64();
      __syncwarp();
      ++li;
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int lg = warp & 3, half = (warp - 2) >> 2;
    const bool fp32 = p.out_fp32 != 0;
    const bool want_stats = !fp32 && p.stats != nullptr;
    float st1[BN / 64], st2[BN / 64];
#pragma unroll
    for (int r = 0; r < BN / 64; ++r) st1[r] = st2[r] = 0.f;
    int cur_n = -1, cur_n0 = -1;
    auto flush = [&]() {
      if (cur_n < 0) return;
#pragma unroll
      for (int r = 0; r < BN / 64; ++r) {
        const int col = cur_n0 + r * 64 + half * 32 + lane;
        if (col < p.ncols) {
          float* dst = p.stats + ((int64_t)cur_n * p.out.C + col) * 2;
          atomicAdd(dst, st1[r]);
          atomicAdd(dst + 1, st2[r]);
        }
        st1[r] = st2[r] = 0.f;
      }
    };
    int li = 0;
    for (int item = it_begin; item < it_end; ++item) {
      PItem it;
      if (!pers_decode(p, pg, item, BN, it)) continue;
      const int bsel = li & 1, use = li >> 1;
      if (want_stats && (it.n != cur_n || it.n0 != cur_n0)) {
        flush();
        cur_n = it.n;
        cur_n0 = it.n0;
      }
      mbar_wait(accf_bar + 8 * bsel, use & 1);
      tc_fence_after();
      if (tsb != nullptr && warp == 2 && lane == 0 && li < 2) tsb[8 + 2 * li] = (unsigned long long)clock64();
      const uint32_t tacc = tmem_base + (uint32_t)(bsel * BN);
      const uint32_t aeb = acce_bar + 8 * bsel;
      if (fp32) {
        if (p.accumulate) pers_epilogue_item<BN, GB_ACT_NONE, true, true>(p, pg, it, tacc, aeb, lg, half, lane, stage, false, st1, st2);
        else pers_epilogue_item<BN, GB_ACT_NONE, true, false>(p, pg, it, tacc, aeb, lg, half, lane, stage, false, st1, st2);
      } else {
        switch (p.act) {
          case GB_ACT_TANH: pers_epilogue_item<BN, GB_ACT_TANH, false, false>(p, pg, it, tacc, aeb, lg, half, lane, stage, want_stats, st1, st2); break;
          case GB_ACT_LEAKY: pers_epilogue_item<BN, GB_ACT_LEAKY, false, false>(p, pg, it, tacc, aeb, lg, half, lane, stage, want_stats, st1, st2); break;
          case GB_ACT_RELU: pers_epilogue_item<BN, GB_ACT_RELU, false, false>(p, pg, it, tacc, aeb, lg, half, lane, stage, want_stats, st1, st2); break;
          default: pers_epilogue_item<BN, GB_ACT_NONE, false, false>(p, pg, it, tacc, aeb, lg, half, lane, stage, want_stats, st1, st2); break;
        }
      }
      if (tsb != nullptr && warp == 2 && lane == 0 && li < 2) tsb[9 + 2 * li] = (unsigned long long)clock64();
      ++li;
    }
    if (want_stats) flush();
  }
  tc_fence_before();
  __syncthreads();
  if (tsb != nullptr && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tsb[12] = (unsigned long long)clock64();
    tsb[13] = gt;
  }
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

unsigned long long* g_timeline = nullptr;
long long g_timeline_ctas = 0;

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    cudaGetLastError();
  }
  return fn;
}

std::mutex g_map_mutex;
std::unordered_map<std::string, CUtensorMap> g_map_cache;

}  // namespace

bool gb_tma_available() { return encode_fn() != nullptr; }

// 5-D activation map {C, W, H, D, N}, box {64, tw, th, 1, 1}
int gb_tma_activation_map(const gb_view& v, int tw, int th, CUtensorMap* out, const int* mul, int c_valid) {
  const int m[3] = {mul ? mul[0] : 1, mul ? mul[1] : 1, mul ? mul[2] : 1};
  std::string key(reinterpret_cast<const char*>(&v), sizeof(gb_view));
  key.append(reinterpret_cast<const char*>(&tw), sizeof(int)).append(reinterpret_cast<const char*>(&th), sizeof(int));
  key.append(reinterpret_cast<const char*>(m), sizeof(m)).append(reinterpret_cast<const char*>(&c_valid), sizeof(int));
  std::lock_guard<std::mutex> lk(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) {
    *out = it->second;
    return 0;
  }
  cuuint64_t dims[5] = {(cuuint64_t)(c_valid > 0 ? c_valid : v.C), (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.D,
                        (cuuint64_t)v.N};
  cuuint64_t strides[4] = {(cuuint64_t)v.sx * 2, (cuuint64_t)v.sy * 2, (cuuint64_t)v.sz * 2, (cuuint64_t)v.sn * 2};
  // with a traversal stride s the box spans n*s coordinates and delivers n elements
  cuuint32_t box[5] = {64, (cuuint32_t)(tw * m[2]), (cuuint32_t)(th * m[1]), (cuuint32_t)m[0], 1};
  cuuint32_t es[5] = {1, (cuuint32_t)m[2], (cuuint32_t)m[1], (cuuint32_t)m[0], 1};
  GB_CHECK(box[1] <= 256 && box[2] <= 256 && box[3] <= 256, "TMA box too large for the gather stride");
  CUresult r = encode_fn()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.ptr, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache[key] = *out;
  return 0;
}

// 5-D activation map {C, W, H, D, N} of a view with 16 or 32 channels per pixel: box {cbox, tw, th, 1, 1}, the swizzle
// whose span is one pixel (SWIZZLE_32B / SWIZZLE_64B) -- igemm_halo_narrow.cu
int gb_tma_activation_map_narrow(const gb_view& v, int cbox, int tw, int th, CUtensorMap* out) {
  GB_CHECK(cbox == 16 || cbox == 32, "narrow activation map: 16 or 32 channels per pixel");
  const int tag = -7;
  std::string key(reinterpret_cast<const char*>(&v), sizeof(gb_view));
  key.append(reinterpret_cast<const char*>(&tw), sizeof(int)).append(reinterpret_cast<const char*>(&th), sizeof(int));
  key.append(reinterpret_cast<const char*>(&cbox), sizeof(int)).append(reinterpret_cast<const char*>(&tag), sizeof(int));
  std::lock_guard<std::mutex> lk(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) {
    *out = it->second;
    return 0;
  }
  cuuint64_t dims[5] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.D, (cuuint64_t)v.N};
  cuuint64_t strides[4] = {(cuuint64_t)v.sx * 2, (cuuint64_t)v.sy * 2, (cuuint64_t)v.sz * 2, (cuuint64_t)v.sn * 2};
  cuuint32_t box[5] = {(cuuint32_t)cbox, (cuuint32_t)tw, (cuuint32_t)th, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = encode_fn()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.ptr, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, cbox == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(narrow activation) failed: %d", (int)r);
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache[key] = *out;
  return 0;
}

// 2-D weight map {kpad, rows} with box {cbox k, bn rows}, cbox = 16 / 32 -> SWIZZLE_32B / SWIZZLE_64B: one tap's
// weights as a K-major tile whose rows are one pixel's channels (igemm_xsplit.cu)
int gb_tma_weight_map_narrow(const void* w, int kpad, int rows, int cbox, int bn, CUtensorMap* out) {
  GB_CHECK(cbox == 16 || cbox == 32, "narrow weight map: 16 or 32 channels per tap");
  struct {
    const void* w;
    int kpad, rows, cbox, bn, tag;
  } k = {w, kpad, rows, cbox, bn, -9};
  std::string key(reinterpret_cast<const char*>(&k), sizeof(k));
  std::lock_guard<std::mutex> lk(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) {
    *out = it->second;
    return 0;
  }
  cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
  cuuint32_t box[2] = {(cuuint32_t)cbox, (cuuint32_t)bn};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, cbox == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(narrow weights) failed: %d", (int)r);
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache[key] = *out;
  return 0;
}

// 2-D weight map {kpad, nclass*npad}, box {64, bn}; all classes of one conv share kpad on this path
int gb_tma_weight_map(const void* w, int kpad, int rows, int bn, CUtensorMap* out) {
  struct {
    const void* w;
    int kpad, rows, bn;
  } k = {w, kpad, rows, bn};
  std::string key(reinterpret_cast<const char*>(&k), sizeof(k));
  std::lock_guard<std::mutex> lk(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) {
    *out = it->second;
    return 0;
  }
  cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)bn};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache[key] = *out;
  return 0;
}

int gb_tma_store_map(const gb_view& v, int tw, int th, const int* mul, int fp32, CUtensorMap* out) {
  const int m[3] = {mul ? mul[0] : 1, mul ? mul[1] : 1, mul ? mul[2] : 1};
  const int esz = fp32 ? 4 : 2;
  std::string key("store", 5);
  key.append(reinterpret_cast<const char*>(&v), sizeof(gb_view));
  key.append(reinterpret_cast<const char*>(&tw), sizeof(int)).append(reinterpret_cast<const char*>(&th), sizeof(int));
  key.append(reinterpret_cast<const char*>(m), sizeof(m)).append(reinterpret_cast<const char*>(&fp32), sizeof(int));
  std::lock_guard<std::mutex> lk(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) {
    *out = it->second;
    return 0;
  }
  cuuint64_t dims[5] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.D, (cuuint64_t)v.N};
  cuuint64_t strides[4] = {(cuuint64_t)v.sx * esz, (cuuint64_t)v.sy * esz, (cuuint64_t)v.sz * esz, (cuuint64_t)v.sn * esz};
  cuuint32_t box[5] = {(cuuint32_t)(128 / esz), (cuuint32_t)(tw * m[2]), (cuuint32_t)(th * m[1]), (cuuint32_t)m[0], 1};
  cuuint32_t es[5] = {1, (cuuint32_t)m[2], (cuuint32_t)m[1], (cuuint32_t)m[0], 1};
  GB_CHECK(box[1] <= 256 && box[2] <= 256 && box[3] <= 256, "TMA store box too large for the output stride");
  CUresult r = encode_fn()(out, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.ptr, dims,
                           strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(output) failed: %d", (int)r);
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache[key] = *out;
  return 0;
}

int gb_tma_f32_matrix_map(const void* ptr, int cols, int rows, int box_rows, CUtensorMap* out) {
  struct {
    const void* w;
    int cols, rows, br, tag;
  } k = {ptr, cols, rows, box_rows, 0x66333264};
  std::string key(reinterpret_cast<const char*>(&k), sizeof(k));
  std::lock_guard<std::mutex> lk(g_map_mutex);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) {
    *out = it->second;
    return 0;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(fp32 matrix) failed: %d", (int)r);
  if (g_map_cache.size() > 4096) g_map_cache.clear();
  g_map_cache[key] = *out;
  return 0;
}

// Pixel-window views (8 pixels x 8 channels read as one 64-channel "pixel", pixel stride 16 B) need a tensor map
// whose pixel stride is smaller than its channel extent; probe once whether the driver encodes it.
extern "C" int gb_tma_window_supported(void) {
  static int cached = -1;
  if (cached >= 0) return cached;
  if (encode_fn() == nullptr) return cached = 0;
  alignas(64) static char dummy[1 << 16];
  CUtensorMap m;
  cuuint64_t dims[5] = {56, 32, 16, 1, 1};
  cuuint64_t strides[4] = {16, 16 * 39, 16 * 39 * 16, 16 * 39 * 16};
  cuuint32_t box[5] = {64, 8, 16, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dummy, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return cached = (r == CUDA_SUCCESS ? 1 : 0);
}

// Bring-up: per-CTA time stamps of the TMA-fed data kernel (buf: device memory, 8 x u64 per CTA; nullptr = off).
extern "C" int gb_debug_timeline(void* buf, long long max_ctas) {
  g_timeline = static_cast<unsigned long long*>(buf);
  g_timeline_ctas = buf != nullptr ? max_ctas : 0;
  return 0;
}

namespace {

template <int BN>
int launch(const gb_conv_params& p, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const TileGeom& tg,
           cudaStream_t st) {
  using C = TCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(igemm_tma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  dim3 grid(tg.ntiles, gb_cdiv(p.ncols, BN), p.nclass);
  // ring depth: never deeper than the K loop; for narrow tiles on grids of two or more waves a ring of <= ~100 KB
  // lets two CTAs share an SM, so that one CTA's prologue / epilogue overlaps the other's main loop (the K loops of
  // the parity-class and first / last layers are only 2..16 blocks long: per-CTA fixed cost dominated them)
  int kb_max = 1;
  for (int c = 0; c < p.nclass; ++c) kb_max = p.cls[c].ntaps * (p.in.C >> 6) > kb_max ? p.cls[c].ntaps * (p.in.C >> 6) : kb_max;
  TileGeom tgl = tg;
  int ns = C::STAGES < kb_max ? C::STAGES : kb_max;
  const int64_t ctas = (int64_t)grid.x * grid.y * grid.z;
  const int shallow = (100 * 1024) / C::STAGE_BYTES;
  if (BN <= 128 && ctas >= 2 * 148 && kb_max <= 18 && shallow >= 3 && g_gb_knobs[8] == 0) ns = ns < shallow ? ns : shallow;
  if (g_gb_knobs[8] >= 2 && ns > g_gb_knobs[8]) ns = g_gb_knobs[8];  // bring-up: force a ring of knob-8 stages
  if (ns < 1) ns = 1;
  tgl.nstages = ns;
  tgl.mode = g_gb_knobs[30];
  tgl.ts = (g_timeline != nullptr && ctas <= g_timeline_ctas) ? g_timeline : nullptr;
  // the TMA-store epilogue stages the whole output tile (bf16: BN * 256 B, fp32: BN * 512 B) in the operand ring
  size_t ring = (size_t)ns * C::STAGE_BYTES;
  if (tgl.store_mode != 0) {
    const bool f32 = tgl.store_mode == 2 || tgl.store_mode == 3 || tgl.store_mode >= 5;
    // (bf16: + 16 KB behind the tile for the partial statistics of the row groups)
    const size_t staging = (size_t)BN * 128 * (f32 ? 4 : 2) + (f32 ? 0 : 16384);
    if (staging > (size_t)C::SMEM - 2048) tgl.store_mode = 0;
    else if (ring < staging) ring = staging;
  }
  gb_klaunch(igemm_tma_kernel<BN>, grid, 256, ring + 2048, st, p, ma, mb, mo, tgl);
  g_gb_knobs[15] = 2;
  GB_LAUNCH_CHECK();
  return 0;
}

int pers_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN>
int launch_pers(const gb_conv_params& p, const CUtensorMap& ma, const CUtensorMap& mb, const TileGeom& tg, cudaStream_t st) {
  using C = PCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(igemm_pers_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  PersGeom pg;
  pg.tg = tg;
  pg.tg.ts = (g_timeline != nullptr && pers_num_sms() <= g_timeline_ctas) ? g_timeline : nullptr;
  pg.ncb = gb_cdiv(p.ncols, BN);
  const int64_t total = (int64_t)tg.ntiles * pg.ncb * p.nclass;
  if (total >= (1ll << 31)) return -1;
  pg.total = (int)total;
  pg.div_tiles = gb_make_fastdiv((uint32_t)tg.ntiles);
  pg.div_ncb = gb_make_fastdiv((uint32_t)pg.ncb);
  int grid = pers_num_sms();
  if (g_gb_knobs[18] > 0 && g_gb_knobs[18] < grid) grid = g_gb_knobs[18];
  if (total < grid) grid = (int)total;
  gb_klaunch(igemm_pers_kernel<BN>, dim3(grid), 320, C::SMEM, st, p, ma, mb, pg);
  g_gb_knobs[15] = 7;
  GB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// Returns -1 when this path does not apply (caller falls back to the gather kernel), 0 on success, >0 on error.
int gb_conv_data_tma(const gb_conv_params& p, cudaStream_t st) {
  if (g_gb_knobs[3] != 0) return -1;                 // knob 3: disable the TMA path
  if (p.in.C % 64 != 0) return -1;
  for (int d = 0; d < 3; ++d)
    if (p.in_mul[d] < 1 || p.in_mul[d] > 4 || (p.in_mul[d] != 1 && g_gb_knobs[0] == 2)) return -1;  // knob 0 = 2: no strided boxes
  if (p.in.pad != 0) return -1;
  if (encode_fn() == nullptr) return -1;
  // every class must use the same padded K (true when every class has the same tap count) -- otherwise the 2-D
  // weight map cannot address class matrices by row offset
  int kpad = p.cls[0].kpad;
  int64_t max_ext[3] = {0, 0, 0};
  for (int c = 0; c < p.nclass; ++c) {
    if (p.cls[c].kpad != kpad || p.cls[c].w_offset != (int64_t)c * p.npad * kpad) return -1;
    if (p.cls[c].ntaps * p.in.C != p.cls[c].kpad && p.cls[c].ntaps * p.in.C > p.cls[c].kpad) return -1;
    int q[3];
    gb_class_extents(p, c, q);
    for (int d = 0; d < 3; ++d) max_ext[d] = q[d] > max_ext[d] ? q[d] : max_ext[d];
  }
  if (max_ext[0] == 0 || max_ext[1] == 0 || max_ext[2] == 0) return 0;
  if ((p.in.sx * 2) % 16 || (p.in.sy * 2) % 16 || (p.in.sz * 2) % 16 || (p.in.sn * 2) % 16) return -1;
  // tile shape: TW x TH <= 128 pixels.  Fewest tiles first (every tile costs a full 128-row MMA pass whatever it
  // covers: 66x66 -> 11x11 patches, 36 tiles, instead of 16x8, 45 tiles), then the fewest unused rows, then the
  // wider tile (longer contiguous stores).  Knob 0 = 1 restricts TW to powers of two with TW*TH = 128.
  int tw = 8, th = 16;
  {
    int64_t best_tiles = -1, best_waste = 0;
    for (int cand = 4; cand <= 128; ++cand) {
      if (g_gb_knobs[0] == 1 && ((cand & (cand - 1)) != 0 || cand < 8)) continue;
      const int ch = BM / cand;
      if (cand * p.in_mul[2] > 256 || ch * p.in_mul[1] > 256) continue;
      const int64_t tiles = (int64_t)gb_cdiv(max_ext[2], cand) * gb_cdiv(max_ext[1], ch);
      const int64_t unused = BM - cand * ch;  // rows of the MMA tile no pixel maps to
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && unused <= best_waste)) {
        best_tiles = tiles;
        best_waste = unused;
        tw = cand;
        th = ch;
      }
    }
  }
  TileGeom tg;
  tg.tw = tw;
  tg.th = th;
  tg.div_tw = gb_make_fastdiv((uint32_t)tw);
  tg.ntx = gb_cdiv(max_ext[2], tw);
  tg.nty = gb_cdiv(max_ext[1], th);
  tg.tiles_x = gb_make_fastdiv((uint32_t)tg.ntx);
  tg.tiles_y = gb_make_fastdiv((uint32_t)tg.nty);
  tg.tiles_z = gb_make_fastdiv((uint32_t)max_ext[0]);
  const int64_t ntiles = (int64_t)tg.ntx * tg.nty * max_ext[0] * p.in.N;
  if (ntiles >= (1ll << 31)) return -1;
  tg.ntiles = (int)ntiles;
  // tile width: the widest BN covering the output channels, then narrower tiles while that lowers the modelled
  // time  waves(148 SMs) x (K blocks x stage bytes + fixed per-CTA cost).  45 row tiles x 256 columns: BN = 128 is
  // one wave of 90 CTAs, BN = 64 would be 180 CTAs = two waves (measured 27.6 us vs 14.9 us for the 128-CTA case).
  int bn = 16;
  while (bn < p.ncols && bn < 256) bn *= 2;
  if (g_gb_knobs[1] > 0) {
    bn = g_gb_knobs[1];
  } else {
    const int64_t kblocks = (int64_t)(kpad / 64);
    int best_bn = bn;
    int64_t best_cost = -1;
    for (int cand = bn; cand >= 64; cand /= 2) {
      const int64_t ctas = ntiles * p.nclass * gb_cdiv(p.ncols, cand);
      const int64_t waves = (ctas + 147) / 148;
      const int64_t cost = waves * (kblocks * (16 + cand / 8) + 500);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_bn = cand;
      }
    }
    bn = best_bn;
  }
  if (bn > p.nclass * p.npad) return -1;  // keep every TMA box inside its tensor
  CUtensorMap ma, mb, mo;
  if (tw * p.in_mul[2] > 256 || th * p.in_mul[1] > 256) return -1;
  if (gb_tma_activation_map(p.in, tw, th, &ma, p.in_mul, p.in_c_valid)) return 1;
  if (gb_tma_weight_map(p.wpacked, kpad, p.nclass * p.npad, bn, &mb)) return 1;
  // persistent kernel (knob 16 = 3; see igemm_pers_kernel): widest column block, coalesced row stores
  if (g_gb_knobs[16] == 3) {
    bool ok = p.out.pad == 0 && (p.out_fp32 || !p.accumulate);
    // knob 17 (with knob 16 = 3): which launches the persistent kernel takes -- bit 0 bf16 destinations, bit 1 fp32
    // destinations with unit output stride, bit 2 fp32 destinations of parity-class (strided-output) launches; 0 = all
    if (g_gb_knobs[17] != 0) {
      const bool strided = p.out_mul[0] != 1 || p.out_mul[1] != 1 || p.out_mul[2] != 1;
      const int cls = !p.out_fp32 ? 1 : (strided ? 4 : 2);
      ok = ok && (g_gb_knobs[17] & cls) != 0;
    }
    const int esz = p.out_fp32 ? 4 : 2;
    ok = ok && ((uintptr_t)p.out.ptr % 16) == 0 && (p.out.sx * esz) % 16 == 0 && (p.out.sy * esz) % 16 == 0 &&
         (p.out.sz * esz) % 16 == 0 && (p.out.sn * esz) % 16 == 0;
    for (int c = 0; c < p.nclass; ++c) ok = ok && p.cls[c].ntaps >= 1;
    int pbn = 64;
    while (pbn < p.ncols && pbn < 256) pbn *= 2;
    if (g_gb_knobs[1] >= 64) pbn = g_gb_knobs[1];
    ok = ok && pbn <= p.nclass * p.npad;
    if (ok) {
      CUtensorMap mbp;
      if (gb_tma_weight_map(p.wpacked, kpad, p.nclass * p.npad, pbn, &mbp)) return 1;
      tg.store_mode = 0;
      tg.mode = 0;
      tg.ts = nullptr;
      tg.nstages = 0;
      int r = -1;
      switch (pbn) {
        case 64: r = launch_pers<64>(p, ma, mbp, tg, st); break;
        case 128: r = launch_pers<128>(p, ma, mbp, tg, st); break;
        case 256: r = launch_pers<256>(p, ma, mbp, tg, st); break;
      }
      if (r >= 0) return r;
    }
  }
  // staged epilogue (knob 29: 0 = default, 1 = per-thread stores, 2 = bulk tensor store, 3 = coalesced warp stores):
  // whole 128-byte channel groups of a plain, 16-byte aligned view
  tg.store_mode = 0;
  tg.mode = 0;
  tg.ts = nullptr;
  mo = ma;  // (unused unless a bulk-store mode is chosen)
  {
    const int esz = p.out_fp32 ? 4 : 2;
    const int inner = 128 / esz;
    // default (measured: profiles/r02d_conv_microbench_epilogues_b8.txt, r02i_*): the staged tile + bulk store for both
    // destination types -- fp32 data gradients 50.1 -> 45.9 us on the residual-block layer; bf16 forward outputs (their
    // statistics summed from the staged tile through a partial-sum scratch) 56 -> 46 / 39 -> 35 / 50 -> 41 us on the
    // stride-2 and transposed layers; whole step 352.9 -> 358.1 img/s
    const int want = g_gb_knobs[29] == 0 ? 2 : g_gb_knobs[29];
    const bool ok = want >= 2 && bn >= 64 && p.out.pad == 0 && p.out.C % inner == 0 &&
                    ((uintptr_t)p.out.ptr % 16) == 0 && (p.out.sx * esz) % 16 == 0 && (p.out.sy * esz) % 16 == 0 &&
                    (p.out.sz * esz) % 16 == 0 && (p.out.sn * esz) % 16 == 0 && (p.out_fp32 || !p.accumulate);
    const bool tma_ok = ok && tw * p.out_mul[2] <= 256 && th * p.out_mul[1] <= 256 && p.out_mul[0] <= 256;
    if (ok && want == 2 && tma_ok) {
      if (gb_tma_store_map(p.out, tw, th, p.out_mul, p.out_fp32, &mo)) return 1;
      tg.store_mode = p.out_fp32 ? (p.accumulate ? 3 : 2) : 1;
    } else if (ok && want == 3) {
      tg.store_mode = p.out_fp32 ? (p.accumulate ? 6 : 5) : 4;
    }
  }
  g_gb_knobs[31] = tg.store_mode;  // READ-BACK (knob 31): epilogue of the last TMA-fed launch
  switch (bn) {
    case 16: return launch<16>(p, ma, mb, mo, tg, st);
    case 32: return launch<32>(p, ma, mb, mo, tg, st);
    case 64: return launch<64>(p, ma, mb, mo, tg, st);
    case 128: return launch<128>(p, ma, mb, mo, tg, st);
    case 256: return launch<256>(p, ma, mb, mo, tg, st);
  }
  return -1;
}
