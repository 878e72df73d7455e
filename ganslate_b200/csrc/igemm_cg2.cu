// CTA-pair (tcgen05 cta_group::2) persistent variant of the TMA-fed implicit-GEMM convolution.
//
// Why (profiles/r01f, r01g, r01h; DESIGN.md "What bounds the convolutions"): a single-CTA tcgen05.mma of shape
// 128 x 256 x 16 reads A (4 KB) and B (8 KB) from shared memory while TMA writes the same 12 KB -- 192 B/clk/SM
// against the 128 B/clk/SM shared memory delivers, so igemm_tma.cu / igemm_pair.cu stop at 53-61 % tensor-pipe
// active whatever is done to their L2 traffic.  Here two CTAs of one TPC form a pair: each stages its OWN 128-pixel
// patch of A and only HALF of the B tile (BN/2 weight rows); one tcgen05.mma.cta_group::2 of shape 256 x BN x 16,
// issued by the leader CTA, multiplies both patches by the whole B tile.  Shared-memory traffic per CTA drops to
// 128 B/clk and L2 -> SM traffic by a third.
//
// The kernel is also persistent: one cluster per TPC walks the work items (class, column block, patch pair) of the
// launch with a double-buffered TMEM accumulator (2 x BN columns), so the epilogue of item i (tcgen05.ld -> bias /
// activation -> global stores, InstanceNorm statistics) overlaps the main loop of item i + 1, and the barrier
// initialisation / TMEM allocation / pipeline fill are paid once per SM instead of once per tile.
//
// Warp roles (320 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (leader CTA only; also owns the
// TMEM allocation), warps 2..9 = epilogue (warp w reads TMEM lanes 32 * (w % 4) ..; the two groups of four warps
// split the BN columns).
//
// Barriers (per CTA unless stated):
//   full[s]       count 1, LEADER's copy used: the leader's producer arrives with expect_tx = bytes of BOTH CTAs;
//                 both producers' TMA loads complete_tx on it (cp.async.bulk.tensor ... .cta_group::2 lets the peer
//                 signal a barrier in the leader's shared memory).
//   empty[s]      count 1, both copies: tcgen05.commit.cta_group::2 ... multicast (mask 0b11) after the MMAs of a
//                 stage -- each producer waits on its own copy.
//   tfull[a]      count 1, both copies: multicast commit after the last MMA of an item -> epilogue warps.
//   tempty[a]     count 16, LEADER's copy used: one arrive per epilogue warp of both CTAs (remote arrive from the peer)
//                 once the warp's tcgen05.ld of accumulator a have completed -> MMA issuer may overwrite it.
//
// STATUS: compiled for sm_100a and reviewed, NOT yet run on a B200 (the GPU budget of round 1 was spent when it was
// written).  Off by default: gb_debug_knob(16, 1) routes eligible gb_conv_data calls to the CTA-pair kernel,
// gb_debug_knob(16, 2) to the single-CTA persistent variant (same structure, cta_group::1 primitives only); the parity
// tests are tests/test_cg2_gpu.py (GB_EXPERIMENTAL=1).
#include <cuda.h>
#include "gb_common.cuh"
#include "gb_geometry.h"
#include "gb_epilogue.cuh"
#include "gb_tma.h"

int gb_tma_weight_map(const void* w, int kpad, int rows, int bn, CUtensorMap* out);

namespace {

constexpr int BM = 128;                 // rows (pixels) per CTA; the pair MMA has M = 256
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int NTHREADS = 320;
constexpr int EPI_WARP0 = 2;            // first epilogue warp
constexpr int EPI_THREADS = 256;
constexpr int MAX_BIAS = 1024;          // output channels whose bias is staged in shared memory
constexpr int MAXS = 8;                 // barrier slots of the smem ring

template <int BN, bool PAIR>
struct CCfg {
  static constexpr int BH_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;  // this CTA's part of the B tile (half in a pair)
  static constexpr int STAGE_BYTES = A_BYTES + BH_BYTES;
  static constexpr int STAGES = (196 * 1024 / STAGE_BYTES) > MAXS ? MAXS : (196 * 1024 / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulators
  static constexpr int SCRATCH_BYTES = 4 * BN * 2 * 4;         // statistics: [4 row warps][BN][2] floats
  static constexpr int TAIL_BYTES = 1024;
  static constexpr int SMEM = STAGES * STAGE_BYTES + SCRATCH_BYTES + TAIL_BYTES + 1024;
};

// PAIR = false: the same persistent, warp-specialised kernel on ONE CTA per SM (cta_group::1, M = 128, whole B tile per
// CTA) -- knob 16 = 2.  It isolates the effect of persistence / epilogue overlap from that of the pair MMA, and is the
// fallback should the pair variant misbehave on hardware.
struct Cg2Geom {
  gb_fastdiv tiles_x, tiles_y, tiles_z;  // tile index -> (n, z, ty, tx)
  gb_fastdiv div_tw;                     // tile row -> (h, w)
  gb_fastdiv div_pairs, div_nb;          // item -> (cls, column block, pair)
  int tw, th;
  int ntiles;                            // row tiles of one class (max over classes)
  int npairs;                            // ceil(ntiles / 2) (PAIR) or ntiles
  int nb;                                // column blocks
  int nitems;                            // nclass * nb * npairs
  int nstages;
  int wd_mclk;                           // watchdog limit in millions of clocks (0 = off)
  int* wd;                               // host-mapped report buffer (8 ints) or NULL
};

// ---------------------------------------------------------------- cluster / cta_group::2 primitives
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP_C:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE_C;\n"
      "bra WAIT_LOOP_C;\n"
      "WAIT_DONE_C:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// Bring-up watchdog (knob 21 = limit in millions of clocks, 0 = off): a wait that exceeds the limit records WHO waited for
// WHAT in a host-mapped buffer (readable after the context died: gb_debug_cg2_watchdog) and traps, so a protocol or
// hardware-semantics mistake costs one failed launch with a diagnosis instead of a hang until the caller's timeout.
struct WaitTag {
  int role;   // 0 producer/empty, 1 mma/tempty, 2 mma/full, 3 epilogue/tfull
  int index;  // stage or accumulator
  int item;
};
template <bool CLUSTER>
__device__ __forceinline__ void wait_wd(uint32_t bar, uint32_t parity, unsigned long long limit, int* wd, uint32_t rank,
                                        const WaitTag& tag) {
  if (limit == 0ull) {
    if constexpr (CLUSTER) mbar_wait_cluster(bar, parity);
    else mbar_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    if constexpr (CLUSTER) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } else {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
    if (done) return;
    if ((unsigned long long)(clock64() - t0) > limit) {
      if (wd != nullptr && atomicCAS(wd, 0, 1) == 0) {  // first reporter wins
        wd[1] = tag.role; wd[2] = tag.index; wd[3] = tag.item; wd[4] = (int)parity; wd[5] = (int)rank;
        wd[6] = (int)blockIdx.x; wd[7] = (int)threadIdx.x;
        __threadfence_system();
      }
      __trap();
    }
  }
}

// TMA loads of a CTA pair: destination = this CTA's shared memory, completion = `cluster_bar` (a shared::cluster
// address, the leader's full barrier)
__device__ __forceinline__ void tma2_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1,
                                             int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 x 16: 128 rows from each CTA] * B[BN x 16: BN/2 rows from each CTA]^T
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at offset `bar` in BOTH CTAs when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

struct Item {
  int cls, n0, x0, y0, z0, n;
  bool valid;   // this CTA's tile exists in class cls
  bool any;     // at least one tile of the pair exists (otherwise every role skips the item)
};

// item index -> this CTA's tile.  Identical arithmetic in every role of both CTAs.
GB_HD bool tile_of(const gb_conv_params& p, const Cg2Geom& g, int cls, uint32_t t, int& x0, int& y0,
                                        int& z0, int& n) {
  int q[3];
  gb_class_extents(p, cls, q);
  uint32_t u = gb_div(t, g.tiles_x);
  x0 = (int)(t - u * g.tiles_x.d) * g.tw;
  t = u;
  u = gb_div(t, g.tiles_y);
  y0 = (int)(t - u * g.tiles_y.d) * g.th;
  t = u;
  u = gb_div(t, g.tiles_z);
  z0 = (int)(t - u * g.tiles_z.d);
  n = (int)u;
  return n < p.in.N && z0 < q[0] && y0 < q[1] && x0 < q[2];
}

template <int BN, bool PAIR>
GB_HD Item decode_item(const gb_conv_params& p, const Cg2Geom& g, uint32_t item, uint32_t rank) {
  Item it;
  uint32_t u = gb_div(item, g.div_pairs);
  const uint32_t pair = item - u * g.div_pairs.d;
  uint32_t v = gb_div(u, g.div_nb);
  it.n0 = (int)(u - v * g.div_nb.d) * BN;
  it.cls = (int)v;
  int x1, y1, z1, n1;
  const uint32_t t_mine = PAIR ? 2 * pair + rank : pair, t_other = 2 * pair + (rank ^ 1u);
  it.valid = t_mine < (uint32_t)g.ntiles && tile_of(p, g, it.cls, t_mine, it.x0, it.y0, it.z0, it.n);
  const bool other = PAIR && t_other < (uint32_t)g.ntiles && tile_of(p, g, it.cls, t_other, x1, y1, z1, n1);
  it.any = it.valid || other;
  if (!it.valid) {  // a tile that does not exist still takes part in the pair MMA: load tile 0, store nothing
    it.x0 = it.y0 = it.z0 = it.n = 0;
  }
  return it;
}

template <int BN, bool PAIR>
__device__ __forceinline__ void cg_body(const gb_conv_params& p, const CUtensorMap& map_a, const CUtensorMap& map_b,
                                        const Cg2Geom& g) {
  using C = CCfg<BN, PAIR>;
  const int STAGES = g.nstages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;   // identical in both CTAs (same kernel, same static smem)
  uint8_t* smem = smem_raw + (base - raw);
  float* scratch = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES);
  uint8_t* tail = smem + STAGES * C::STAGE_BYTES + C::SCRATCH_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);  // full[8], empty[8], tfull[2], tempty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 8 * (2 * MAXS + 4));
  int8_t* taps_s = reinterpret_cast<int8_t*>(tail + 256);  // GB_MAX_TAPS x 4 bytes
  __shared__ float bias_s[MAX_BIAS];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const uint32_t cluster_id = PAIR ? blockIdx.x >> 1 : blockIdx.x;
  const uint32_t nclusters = PAIR ? gridDim.x >> 1 : gridDim.x;
  const int chunks = p.in.C >> 6;
  const unsigned long long wd_limit = (unsigned long long)g.wd_mclk * 1000000ull;

  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = smem_u32(bars + MAXS);
  const uint32_t tfull_bar = smem_u32(bars + 2 * MAXS);
  const uint32_t tempty_bar = smem_u32(bars + 2 * MAXS + 2);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, PAIR ? 16 : 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc2<C::TMEM_COLS>(smem_u32(tmem_slot));
    else tmem_alloc<C::TMEM_COLS>(smem_u32(tmem_slot));
  }
  for (int i = tid; i < GB_MAX_TAPS; i += NTHREADS)
    *reinterpret_cast<uint32_t*>(taps_s + 4 * i) = *reinterpret_cast<const uint32_t*>(p.taps[i]);
  for (int i = tid; i < MAX_BIAS; i += NTHREADS) bias_s[i] = (p.bias != nullptr && i < p.ncols) ? p.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // barrier inits / TMEM allocations of both CTAs precede any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one lane per CTA)
    if (lane == 0) {
      const uint32_t leader_full = PAIR ? map_to_cta(full_bar, 0) : full_bar;
      const uint32_t tx_bytes = (PAIR ? 2u : 1u) * (uint32_t)(g.tw * g.th * 128 + C::BH_BYTES);
      int s = 0, round = 0;
      for (uint32_t item = cluster_id; item < (uint32_t)g.nitems; item += nclusters) {
        const Item it = decode_item<BN, PAIR>(p, g, item, rank);
        if (!it.any) continue;
        const gb_conv_class& cc = p.cls[it.cls];
        for (int tl = 0; tl < cc.ntaps; ++tl) {
          const int8_t* tp = taps_s + 4 * (cc.tap_begin + tl);
          const int dz = tp[0], dy = tp[1], dx = tp[2];
          for (int c = 0; c < chunks; ++c) {
            if (round > 0) wait_wd<false>(empty_bar + 8 * s, (round - 1) & 1, wd_limit, g.wd, rank, WaitTag{0, s, (int)item});
            const uint32_t a_s = base + s * C::STAGE_BYTES;
            const uint32_t b_s = a_s + A_BYTES;
            if (leader) mbar_expect_tx(full_bar + 8 * s, tx_bytes);
            if constexpr (PAIR) {
              tma2_load_5d(a_s, &map_a, leader_full + 8 * s, c * 64, it.x0 * p.in_mul[2] + dx, it.y0 * p.in_mul[1] + dy,
                           it.z0 * p.in_mul[0] + dz, it.n);
              tma2_load_2d(b_s, &map_b, leader_full + 8 * s, tl * p.in.C + c * 64,
                           it.cls * p.npad + it.n0 + (int)rank * (BN / 2));
            } else {
              tma_load_5d(a_s, &map_a, leader_full + 8 * s, c * 64, it.x0 * p.in_mul[2] + dx, it.y0 * p.in_mul[1] + dy,
                          it.z0 * p.in_mul[0] + dz, it.n);
              tma_load_2d(b_s, &map_b, leader_full + 8 * s, tl * p.in.C + c * 64, it.cls * p.npad + it.n0);
            }
            if (++s == STAGES) {
              s = 0;
              ++round;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one lane)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = PAIR ? make_idesc_bf16_m256(BN) : make_idesc_bf16(BN, 0, 0);
      int s = 0, round = 0;
      uint32_t acc_it = 0;  // items issued so far: accumulator = acc_it & 1, its use count = acc_it >> 1
      for (uint32_t item = cluster_id; item < (uint32_t)g.nitems; item += nclusters) {
        const Item it = decode_item<BN, PAIR>(p, g, item, rank);
        if (!it.any) continue;
        const gb_conv_class& cc = p.cls[it.cls];
        const int KB = cc.ntaps * chunks;
        const uint32_t a = acc_it & 1u, use = acc_it >> 1;
        if (use > 0) {  // the epilogue warps of both CTAs have drained the previous use of this accumulator
          wait_wd<PAIR>(tempty_bar + 8 * a, (use - 1) & 1, wd_limit, g.wd, rank, WaitTag{1, (int)a, (int)item});
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int kb = 0; kb < KB; ++kb) {
          wait_wd<false>(full_bar + 8 * s, round & 1, wd_limit, g.wd, rank, WaitTag{2, s, (int)item});
          tc_fence_after();
          const uint32_t a_s = base + s * C::STAGE_BYTES;
          const uint32_t b_s = a_s + A_BYTES;
          const uint64_t adesc = make_smem_desc(a_s, 16, 1024);
          const uint64_t bdesc = make_smem_desc(b_s, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if constexpr (PAIR) umma2_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
            else umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
          }
          if constexpr (PAIR) umma2_commit(empty_bar + 8 * s);
          else umma_commit(empty_bar + 8 * s);
          if (++s == STAGES) {
            s = 0;
            ++round;
          }
        }
        if constexpr (PAIR) umma2_commit(tfull_bar + 8 * a);
        else umma_commit(tfull_bar + 8 * a);
        ++acc_it;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    const int lg = warp & 3;                    // TMEM lane group this warp may read
    const int half = (warp - EPI_WARP0) >> 2;   // which half of the BN columns
    const int etid = tid - EPI_WARP0 * 32;      // 0..255
    const uint32_t leader_tempty = PAIR ? map_to_cta(tempty_bar, 0) : tempty_bar;
    const bool want_stats = p.stats != nullptr && !p.out_fp32;
    __nv_bfloat16* optr = reinterpret_cast<__nv_bfloat16*>(p.out.ptr);
    uint32_t acc_it = 0;
    for (uint32_t item = cluster_id; item < (uint32_t)g.nitems; item += nclusters) {
      const Item it = decode_item<BN, PAIR>(p, g, item, rank);
      if (!it.any) continue;
      const gb_conv_class& cc = p.cls[it.cls];
      const uint32_t a = acc_it & 1u, use = acc_it >> 1;
      ++acc_it;
      int q[3];
      gb_class_extents(p, it.cls, q);
      const int row = lg * 32 + lane;
      const int h = (int)gb_div((uint32_t)row, g.div_tw), w = row - h * g.tw;
      const int qy = it.y0 + h, qx = it.x0 + w;
      const bool row_ok = it.valid && h < g.th && qy < q[1] && qx < q[2];
      int64_t ooff = 0;
      if (row_ok)
        ooff = gb_pix_offset(p.out, it.n, it.z0 * p.out_mul[0] + cc.off[0], qy * p.out_mul[1] + cc.off[1],
                             qx * p.out_mul[2] + cc.off[2]);
      wait_wd<false>(tfull_bar + 8 * a, use & 1, wd_limit, g.wd, rank, WaitTag{3, (int)a, (int)item});
      tc_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(lg * 32) << 16) + a * BN;
      constexpr int CH = 32;
      const int cbeg = half * (BN / 2);
#pragma unroll 1
      for (int c0 = cbeg; c0 < cbeg + BN / 2; c0 += CH) {
        uint32_t acc[CH];
        tmem_ld32(t_acc + (uint32_t)c0, acc);
        tmem_ld_wait();
        if (c0 + CH >= cbeg + BN / 2) {
          // last TMEM read of this warp for this accumulator: hand it back to the MMA issuer before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(leader_tempty + 8 * a);
            else mbar_arrive(leader_tempty + 8 * a);
          }
        }
        float sv[CH];
#pragma unroll
        for (int gq = 0; gq < CH / 8; ++gq) {
          const int col = it.n0 + c0 + gq * 8;
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int bc = col + e < MAX_BIAS ? col + e : MAX_BIAS - 1;
            float t = __uint_as_float(acc[gq * 8 + e]) + bias_s[bc];
            if (p.act == GB_ACT_TANH) t = tanhf(t);
            else if (p.act == GB_ACT_LEAKY) t = t > 0.f ? t : t * p.act_slope;
            else if (p.act == GB_ACT_RELU) t = fmaxf(t, 0.f);
            v[e] = t;
          }
          const bool col_ok = row_ok && col < p.out.C;
          if (p.out_fp32) {
            if (col_ok) {
              float4* o32 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out.ptr) + ooff + col);
              float4 x = make_float4(v[0], v[1], v[2], v[3]), y = make_float4(v[4], v[5], v[6], v[7]);
              if (p.accumulate) {
                const float4 pa = o32[0], pb = o32[1];
                x.x += pa.x; x.y += pa.y; x.z += pa.z; x.w += pa.w;
                y.x += pb.x; y.y += pb.y; y.z += pb.z; y.w += pb.w;
              }
              o32[0] = x;
              o32[1] = y;
            }
          } else {
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]);
            o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]);
            o.w = pack_bf16x2(v[6], v[7]);
            if (col_ok) *reinterpret_cast<uint4*>(optr + ooff + col) = o;
            if (want_stats) {
              float2 f;
              f = unpack_bf16x2(o.x); sv[gq * 8 + 0] = f.x; sv[gq * 8 + 1] = f.y;
              f = unpack_bf16x2(o.y); sv[gq * 8 + 2] = f.x; sv[gq * 8 + 3] = f.y;
              f = unpack_bf16x2(o.z); sv[gq * 8 + 4] = f.x; sv[gq * 8 + 5] = f.y;
              f = unpack_bf16x2(o.w); sv[gq * 8 + 6] = f.x; sv[gq * 8 + 7] = f.y;
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (!row_ok || col + e >= p.ncols) sv[gq * 8 + e] = 0.f;
            }
          }
        }
        if (want_stats) {  // uniform over the launch
          float sq[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) sq[i] = sv[i] * sv[i];
          const float s1 = gb_warp_colsum<CH>(sv, lane);
          const float s2 = gb_warp_colsum<CH>(sq, lane);
          scratch[(lg * BN + c0 + lane) * 2 + 0] = s1;
          scratch[(lg * BN + c0 + lane) * 2 + 1] = s2;
        }
      }
      if (want_stats) {
        // a tile lies inside ONE image: sum the four row warps in shared memory, one atomic per column and moment
        epi_bar_sync();
        if (it.valid) {
          for (int i = etid; i < BN; i += EPI_THREADS) {
            const int col = it.n0 + i;
            if (col < p.ncols) {
              const float s1 = scratch[i * 2] + scratch[(BN + i) * 2] + scratch[(2 * BN + i) * 2] + scratch[(3 * BN + i) * 2];
              const float s2 = scratch[i * 2 + 1] + scratch[(BN + i) * 2 + 1] + scratch[(2 * BN + i) * 2 + 1] +
                               scratch[(3 * BN + i) * 2 + 1];
              float* dst = p.stats + ((int64_t)it.n * p.out.C + col) * 2;
              atomicAdd(dst, s1);
              atomicAdd(dst + 1, s2);
            }
          }
        }
        epi_bar_sync();  // scratch is rewritten by the next item
      }
    }
  }

  // -------------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // no CTA of the pair leaves (or frees TMEM) while the other may still signal it
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc2<C::TMEM_COLS>(tmem_base);
    else tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
igemm_cg2_kernel(const __grid_constant__ gb_conv_params p, const __grid_constant__ CUtensorMap map_a,
                 const __grid_constant__ CUtensorMap map_b, const __grid_constant__ Cg2Geom g) {
  gb_pdl_enter();
  cg_body<BN, true>(p, map_a, map_b, g);
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
igemm_persist_kernel(const __grid_constant__ gb_conv_params p, const __grid_constant__ CUtensorMap map_a,
                     const __grid_constant__ CUtensorMap map_b, const __grid_constant__ Cg2Geom g) {
  gb_pdl_enter();
  cg_body<BN, false>(p, map_a, map_b, g);
}

int* g_wd_host = nullptr;   // host-mapped watchdog report (8 ints), allocated on first use
int* g_wd_dev = nullptr;

void watchdog_setup(Cg2Geom& g) {
  g.wd_mclk = g_gb_knobs[21] > 0 ? g_gb_knobs[21] : 0;
  g.wd = nullptr;
  if (g.wd_mclk == 0) return;
  if (g_wd_host == nullptr) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&g_wd_host), 8 * sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_wd_dev), g_wd_host, 0) != cudaSuccess) {
      cudaGetLastError();
      g_wd_host = g_wd_dev = nullptr;
      g.wd_mclk = 0;
      return;
    }
    for (int i = 0; i < 8; ++i) g_wd_host[i] = 0;
  }
  g.wd = g_wd_dev;
}

template <int BN, bool PAIR>
int launch(const gb_conv_params& p, const CUtensorMap& ma, const CUtensorMap& mb, Cg2Geom g, cudaStream_t st) {
  using C = CCfg<BN, PAIR>;
  watchdog_setup(g);
  static bool attr_set = false;
  static int max_groups = 0;  // co-resident clusters (PAIR) or CTAs
  if (!attr_set) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    if constexpr (PAIR) {
      GB_CUDA(cudaFuncSetAttribute(igemm_cg2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(sms, 1, 1);
      cfg.blockDim = dim3(NTHREADS, 1, 1);
      cfg.dynamicSmemBytes = C::SMEM;
      int nc = 0;
      // clusters never wait for each other, so over-subscription is harmless: fall back to one pair per TPC
      if (cudaOccupancyMaxActiveClusters(&nc, igemm_cg2_kernel<BN>, &cfg) != cudaSuccess || nc <= 0) nc = sms / 2;
      cudaGetLastError();
      max_groups = nc;
    } else {
      GB_CUDA(cudaFuncSetAttribute(igemm_persist_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
      max_groups = sms;  // one CTA per SM (the smem ring allows no second one)
    }
    attr_set = true;
  }
  if (max_groups <= 0) return -1;
  // the ring may be deeper than one item's K loop: the producer runs ahead into the next item
  g.nstages = g_gb_knobs[17] > 0 && g_gb_knobs[17] <= C::STAGES ? g_gb_knobs[17] : C::STAGES;
  int ngroups = max_groups < g.nitems ? max_groups : g.nitems;
  if (g_gb_knobs[18] > 0 && g_gb_knobs[18] < ngroups) ngroups = g_gb_knobs[18];
  if constexpr (PAIR) gb_klaunch(igemm_cg2_kernel<BN>, dim3(2 * ngroups, 1, 1), NTHREADS, C::SMEM, st, p, ma, mb, g);
  else gb_klaunch(igemm_persist_kernel<BN>, dim3(ngroups, 1, 1), NTHREADS, C::SMEM, st, p, ma, mb, g);
  g_gb_knobs[15] = PAIR ? 5 : 6;
  g_gb_knobs[19] += 1;  // launches served here (tests read and reset it)
  GB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// Geometry of a launch: -1 = path does not apply, 0 = nothing to do, 1 = geometry filled.
static int cg2_geometry(const gb_conv_params& p, bool pair, Cg2Geom* gout, int* bn_out, int* kpad_out) {
  if (p.in.C % 64 != 0 || p.in.pad != 0) return -1;
  if (p.ncols > MAX_BIAS || p.ncols < 33) return -1;  // narrow outputs stay on the one-tile-per-CTA kernels
  for (int d = 0; d < 3; ++d)
    if (p.in_mul[d] < 1 || p.in_mul[d] > 4) return -1;
  const int kpad = p.cls[0].kpad;
  int64_t max_ext[3] = {0, 0, 0};
  for (int c = 0; c < p.nclass; ++c) {
    if (p.cls[c].kpad != kpad || p.cls[c].w_offset != (int64_t)c * p.npad * kpad) return -1;
    if (p.cls[c].ntaps < 1 || p.cls[c].ntaps * p.in.C > p.cls[c].kpad) return -1;
    int q[3];
    gb_class_extents(p, c, q);
    for (int d = 0; d < 3; ++d) max_ext[d] = q[d] > max_ext[d] ? q[d] : max_ext[d];
  }
  if (max_ext[0] == 0 || max_ext[1] == 0 || max_ext[2] == 0) return 0;
  if ((p.in.sx * 2) % 16 || (p.in.sy * 2) % 16 || (p.in.sz * 2) % 16 || (p.in.sn * 2) % 16) return -1;
  // patch shape: as igemm_tma.cu (fewest tiles, then fewest unused rows, then the wider patch)
  int tw = 8, th = 16;
  {
    int64_t best_tiles = -1, best_waste = 0;
    for (int cand = 4; cand <= 128; ++cand) {
      const int ch = BM / cand;
      if (cand * p.in_mul[2] > 256 || ch * p.in_mul[1] > 256) continue;
      const int64_t tiles = (int64_t)gb_cdiv(max_ext[2], cand) * gb_cdiv(max_ext[1], ch);
      const int64_t unused = BM - cand * ch;
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && unused <= best_waste)) {
        best_tiles = tiles;
        best_waste = unused;
        tw = cand;
        th = ch;
      }
    }
  }
  Cg2Geom g;
  g.tw = tw;
  g.th = th;
  g.div_tw = gb_make_fastdiv((uint32_t)tw);
  const int ntx = gb_cdiv(max_ext[2], tw), nty = gb_cdiv(max_ext[1], th);
  g.tiles_x = gb_make_fastdiv((uint32_t)ntx);
  g.tiles_y = gb_make_fastdiv((uint32_t)nty);
  g.tiles_z = gb_make_fastdiv((uint32_t)max_ext[0]);
  const int64_t ntiles = (int64_t)ntx * nty * max_ext[0] * p.in.N;
  if (ntiles >= (1ll << 30)) return -1;
  g.ntiles = (int)ntiles;
  g.npairs = pair ? (int)((ntiles + 1) / 2) : (int)ntiles;
  // column block: the widest MMA N covering the output channels (<= 256); knob 1 overrides
  int bn = 64;
  while (bn < p.ncols && bn < 256) bn *= 2;
  if (g_gb_knobs[1] >= 64) bn = g_gb_knobs[1];
  const int box_rows = pair ? bn / 2 : bn;
  if (box_rows > p.nclass * p.npad) return -1;  // keep the weight box inside its tensor
  g.nb = gb_cdiv(p.ncols, bn);
  const int64_t nitems = (int64_t)p.nclass * g.nb * g.npairs;
  if (nitems >= (1ll << 31)) return -1;
  g.nitems = (int)nitems;
  g.div_pairs = gb_make_fastdiv((uint32_t)g.npairs);
  g.div_nb = gb_make_fastdiv((uint32_t)g.nb);
  g.nstages = 0;
  g.wd_mclk = 0;
  g.wd = nullptr;
  *gout = g;
  *bn_out = bn;
  *kpad_out = kpad;
  return 1;
}

// Returns -1 when this path does not apply or is switched off (knob 16: 0 = off, 1 = CTA pair, 2 = persistent single
// CTA), 0 on success, >0 on error.
int gb_conv_data_cg2(const gb_conv_params& p, cudaStream_t st) {
  const int mode = g_gb_knobs[16];
  if (mode == 0 || mode >= 3 || g_gb_knobs[3] != 0) return -1;   // (3 = igemm_pers_kernel in igemm_tma.cu)
  const bool pair = mode == 1;
  if (!gb_tma_available()) return -1;
  Cg2Geom g;
  int bn = 0, kpad = 0;
  const int r = cg2_geometry(p, pair, &g, &bn, &kpad);
  if (r <= 0) return r;
  CUtensorMap ma, mb;
  if (gb_tma_activation_map(p.in, g.tw, g.th, &ma, p.in_mul, p.in_c_valid)) return 1;
  if (gb_tma_weight_map(p.wpacked, kpad, p.nclass * p.npad, pair ? bn / 2 : bn, &mb)) return 1;
  if (pair) {
    switch (bn) {
      case 64: return launch<64, true>(p, ma, mb, g, st);
      case 128: return launch<128, true>(p, ma, mb, g, st);
      case 256: return launch<256, true>(p, ma, mb, g, st);
    }
  } else {
    switch (bn) {
      case 64: return launch<64, false>(p, ma, mb, g, st);
      case 128: return launch<128, false>(p, ma, mb, g, st);
      case 256: return launch<256, false>(p, ma, mb, g, st);
    }
  }
  return -1;
}

// The watchdog's report after a trapped launch (knob 21): out[0] = 1 when a wait timed out, then {role (0 producer waits
// for an empty stage, 1 MMA issuer waits for a drained accumulator, 2 MMA issuer waits for a full stage, 3 epilogue
// waits for a complete accumulator), stage / accumulator index, item, parity, CTA rank in the pair, blockIdx.x,
// threadIdx.x}.  Reads host memory only, so it works after the CUDA context has been lost.  Returns out[0].
extern "C" int gb_debug_cg2_watchdog(int32_t* out) {
  for (int i = 0; i < 8; ++i) out[i] = g_wd_host ? g_wd_host[i] : 0;
  return out[0];
}

// Host replay of the persistent kernels' work decomposition (no device needed): for item i and CTA rank r of the pair
// (rank 0 only in mode 2) writes 8 ints {cls, n0, x0, y0, z0, n, valid, any} at out[(i * ranks + r) * 8], with the SAME
// decode_item code the kernels run, and the patch shape / column block / item count into info[0..5] =
// {tw, th, bn, nitems, ranks, ntiles}.  Returns the number of items (0 when out is too small: call with out = NULL
// first), -1 when the launch would not take this path.  CPU tests check that every output pixel of every class and
// column block is produced exactly once.
extern "C" int gb_debug_cg2_plan(const gb_conv_params* pp, int mode, int32_t* info, int32_t* out, int64_t out_ints) {
  if (pp == nullptr || info == nullptr || (mode != 1 && mode != 2)) return -1;
  const bool pair = mode == 1;
  Cg2Geom g;
  int bn = 0, kpad = 0;
  const int r = cg2_geometry(*pp, pair, &g, &bn, &kpad);
  if (r < 0) return -1;
  const int ranks = pair ? 2 : 1;
  if (r == 0) {
    info[0] = info[1] = info[2] = info[3] = info[5] = 0;
    info[4] = ranks;
    return 0;
  }
  info[0] = g.tw; info[1] = g.th; info[2] = bn; info[3] = g.nitems; info[4] = ranks; info[5] = g.ntiles;
  if (out == nullptr || out_ints < (int64_t)g.nitems * ranks * 8) return 0;
  for (int i = 0; i < g.nitems; ++i)
    for (int rk = 0; rk < ranks; ++rk) {
      Item it;
      switch (bn) {  // n0 depends on the column-block width only through BN
        case 64: it = pair ? decode_item<64, true>(*pp, g, (uint32_t)i, (uint32_t)rk) : decode_item<64, false>(*pp, g, (uint32_t)i, 0u); break;
        case 128: it = pair ? decode_item<128, true>(*pp, g, (uint32_t)i, (uint32_t)rk) : decode_item<128, false>(*pp, g, (uint32_t)i, 0u); break;
        default: it = pair ? decode_item<256, true>(*pp, g, (uint32_t)i, (uint32_t)rk) : decode_item<256, false>(*pp, g, (uint32_t)i, 0u); break;
      }
      int32_t* o = out + ((int64_t)i * ranks + rk) * 8;
      o[0] = it.cls; o[1] = it.n0; o[2] = it.x0; o[3] = it.y0; o[4] = it.z0; o[5] = it.n; o[6] = it.valid ? 1 : 0;
      o[7] = it.any ? 1 : 0;
    }
  return g.nitems;
}
