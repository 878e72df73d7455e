// Implicit-GEMM convolution (forward / data-gradient / transposed) on tcgen05 tensor cores.
//
//   D[128 x BN] (TMEM, fp32)  +=  A[128 x 64] (gathered activations, K-major, smem)
//                                * B[BN x 64]^T (packed weights, K-major, smem)
//
// GEMM rows enumerate the q-grid of one parity class, the K index is (tap, channel).  Operand tiles are
// gathered in 16-byte chunks with cp.async (zero-fill outside the image) straight into the 128B-swizzled
// layout the UMMA descriptors expect, so stride, zero padding, transposed convolution (as parity classes)
// and ragged channel counts all share this one kernel.  Reflection padding is materialised by the producer
// of the activation (see instnorm.cu), never here.
//
// Warp roles (256 threads): warps 0-3 gather, warp 4 lane 0 issues tcgen05.mma, all 8 warps run the epilogue
// (TMEM -> registers -> +bias / activation -> bf16 -> global).
#include "gb_common.cuh"
#include "gb_geometry.h"
#include "gb_epilogue.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int LAG = 2;                // cp.async groups in flight before the oldest is published

template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 1024 /*barriers, taps, bias*/;
  static constexpr int MIN_CTAS = (BN == 256) ? 1 : 2;
};

struct ClassDivs {
  gb_fastdiv f[GB_MAX_CLASSES][3];  // divide by the (z, y, x) q-grid extents of each class
};

template <int BN>
__global__ void __launch_bounds__(256, Cfg<BN>::MIN_CTAS) igemm_data_kernel(const __grid_constant__ gb_conv_params p,
                                                                            const __grid_constant__ ClassDivs divs) {
  gb_pdl_enter();
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  // tail region after the stages
  uint8_t* tail = smem + STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);             // full[STAGES], empty[STAGES], accum
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 128);
  int8_t* taps_s = reinterpret_cast<int8_t*>(tail + 192);          // up to 128*4 = 512 B
  // bias lives right after: tail + 704 .. needs BN*4 <= 1024 -> put it in its own static array instead
  __shared__ float bias_s[BN];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cls = blockIdx.z;
  const gb_conv_class& cc = p.cls[cls];
  int q[3];
  gb_class_extents(p, cls, q);
  const int64_t Mc = (int64_t)p.in.N * q[0] * q[1] * q[2];
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  if (m0 >= Mc) return;
  const int n0 = blockIdx.y * BN;
  const int KB = (cc.ntaps * p.in.C + BK - 1) / BK;  // kpad is only the row pitch of the packed class matrix

  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = smem_u32(bars + STAGES);
  const uint32_t accum_bar = smem_u32(bars + 2 * STAGES);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 4);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<C::TMEM_COLS>(smem_u32(tmem_slot));
  for (int i = tid; i < cc.ntaps; i += 256)
    *reinterpret_cast<uint32_t*>(taps_s + 4 * i) = *reinterpret_cast<const uint32_t*>(p.taps[cc.tap_begin + i]);
  for (int i = tid; i < BN; i += 256) bias_s[i] = (p.bias != nullptr && n0 + i < p.ncols) ? p.bias[n0 + i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ gather producers
    // Thread (r0, j) fills 16-byte chunk j of rows r0 + 16*i.  The warp is issue-latency bound (one warp per
    // scheduler), so everything that does not change from stage to stage is hoisted: row decode, the swizzled
    // smem offsets, and -- when a stage never straddles two taps (C % 64 == 0) -- the per-tap bounds mask and
    // source offsets, which are recomputed only when the tap changes.
    const int j = tid & 7;    // 16-byte chunk inside the 128-byte row
    const int r0 = tid >> 3;  // first row handled by this thread (rows r0 + 16*i)
    const int C8 = p.in.C >> 3;
    int rbase[8], ryx[8], rz[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t m = m0 + r0 + 16 * i;
      if (m < Mc) {
        gb_row r = gb_decode_row_fast((uint32_t)m, divs.f[cls]);
        const int gz = r.qz * p.in_mul[0], gy = r.qy * p.in_mul[1], gx = r.qx * p.in_mul[2];
        rbase[i] = (int)gb_pix_offset(p.in, r.n, gz, gy, gx);
        ryx[i] = (gy << 16) | (gx & 0xFFFF);
        rz[i] = gz;
      } else {
        rbase[i] = 0;
        ryx[i] = (int)0x80008000;  // y = x = -32768: always out of bounds
        rz[i] = 0;
      }
    }
    const __nv_bfloat16* in_ptr = reinterpret_cast<const __nv_bfloat16*>(p.in.ptr);
    const __nv_bfloat16* w_ptr = reinterpret_cast<const __nv_bfloat16*>(p.wpacked) + cc.w_offset;
    // swizzled destination of this thread's chunk: (row & 7) == (r0 & 7) for all its rows
    const uint32_t dst0 = (uint32_t)r0 * 128u + (uint32_t)((j ^ (r0 & 7)) << 4);
    // weight rows
    int boff[BN / 16 > 0 ? BN / 16 : 1];
    uint32_t bmask = 0;
#pragma unroll
    for (int i = 0; i < BN / 16; ++i) {
      const int n = n0 + r0 + 16 * i;
      boff[i] = n * cc.kpad + j * 8;
      if (n < p.npad) bmask |= 1u << i;
    }
    const bool fast = (C8 & 7) == 0;   // a 64-wide K block lies inside one tap
    const int chunks_per_tap = C8 >> 3;
    int tap = -1, cc_in_tap = 0;       // fast path state
    uint32_t amask = 0;
    int aoff[8];
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % STAGES;
      const int it = kb / STAGES;
      if (it > 0) mbar_wait(empty_bar + 8 * s, (it - 1) & 1);
      const uint32_t a_s = base + s * C::STAGE_BYTES + dst0;
      const uint32_t b_s = a_s + A_BYTES;
      if (fast) {
        if (tap < 0 || cc_in_tap == chunks_per_tap) {
          ++tap;
          cc_in_tap = 0;
          const bool tap_ok = tap < cc.ntaps;
          int dz = 0, dy = 0, dx = 0;
          if (tap_ok) {
            dz = taps_s[4 * tap + 0];
            dy = taps_s[4 * tap + 1];
            dx = taps_s[4 * tap + 2];
          }
          const int toff = (int)(dz * p.in.sz + dy * p.in.sy + dx * p.in.sx) + j * 8;
          amask = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int z = rz[i] + dz, y = (ryx[i] >> 16) + dy, x = (int)(short)(ryx[i] & 0xFFFF) + dx;
            if (tap_ok && gb_in_bounds(p.in, z, y, x)) amask |= 1u << i;
            aoff[i] = rbase[i] + toff;
          }
        }
        const int coff = cc_in_tap * 64;
        ++cc_in_tap;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = (amask >> i) & 1u;
          cp_async16(a_s + i * 2048, ok ? in_ptr + (aoff[i] + coff) : in_ptr, ok);
        }
      } else {
        const int k8 = kb * 8 + j;
        const int tl = k8 / C8;
        const int c8 = k8 - tl * C8;
        const bool tap_ok = tl < cc.ntaps;
        int dz = 0, dy = 0, dx = 0;
        if (tap_ok) {
          dz = taps_s[4 * tl + 0];
          dy = taps_s[4 * tl + 1];
          dx = taps_s[4 * tl + 2];
        }
        const int toff = (int)(dz * p.in.sz + dy * p.in.sy + dx * p.in.sx) + c8 * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int z = rz[i] + dz, y = (ryx[i] >> 16) + dy, x = (int)(short)(ryx[i] & 0xFFFF) + dx;
          const bool ok = tap_ok && gb_in_bounds(p.in, z, y, x);
          cp_async16(a_s + i * 2048, ok ? in_ptr + (rbase[i] + toff) : in_ptr, ok);
        }
      }
      const int kcol = kb * BK;
#pragma unroll
      for (int i = 0; i < BN / 16; ++i) {
        const bool ok = (bmask >> i) & 1u;
        cp_async16(b_s + i * 2048, ok ? w_ptr + (boff[i] + kcol) : w_ptr, ok);
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
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(BN, 0, 0);
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % STAGES;
      const int it = kb / STAGES;
      mbar_wait(full_bar + 8 * s, it & 1);
      tc_fence_after();
      if (lane == 0) {
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
    }
    if (lane == 0 && KB > 0) umma_commit(accum_bar);
    __syncwarp();
  }

  // -------------------------------------------------------------------- epilogue (all warps)
  if (KB > 0) {
    mbar_wait(accum_bar, 0);
    tc_fence_after();
  }
  {
    const int row = (warp & 3) * 32 + lane;
    const int64_t m = m0 + row;
    const bool row_ok = m < Mc;
    int64_t ooff = 0;
    int row_n = 0;
    if (row_ok) {
      gb_row r = gb_decode_row_fast((uint32_t)m, divs.f[cls]);
      row_n = r.n;
      ooff = gb_pix_offset(p.out, r.n, r.qz * p.out_mul[0] + cc.off[0], r.qy * p.out_mul[1] + cc.off[1],
                           r.qx * p.out_mul[2] + cc.off[2]);
    }
    gb_conv_epilogue<BN>(p, tmem_base, warp, lane, KB > 0, row_ok, ooff, n0, bias_s, row_n);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int BN>
int launch(const gb_conv_params& p, const ClassDivs& divs, int64_t max_mc, cudaStream_t st) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(igemm_data_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  dim3 grid(gb_cdiv(max_mc, BM), gb_cdiv(p.ncols, BN), p.nclass);
  gb_klaunch(igemm_data_kernel<BN>, grid, 256, C::SMEM, st, p, divs);
  g_gb_knobs[15] = 1;
  GB_LAUNCH_CHECK();
  return 0;
}

int64_t view_max_offset(const gb_view& v) {
  return (int64_t)(v.N - 1) * v.sn + (int64_t)(v.D - 1) * v.sz + (int64_t)(v.H - 1 + v.pad) * v.sy +
         (int64_t)(v.W - 1 + v.pad) * v.sx + v.C;
}

}  // namespace

int gb_conv_data_tma(const gb_conv_params& p, cudaStream_t st);   // igemm_tma.cu: -1 = not applicable
int gb_conv_data_halo(const gb_conv_params& p, cudaStream_t st);  // igemm_halo.cu: -1 = not applicable
int gb_conv_data_halo_narrow(const gb_conv_params& p, cudaStream_t st);  // igemm_halo_narrow.cu: -1 = not applicable
int gb_conv_data_xsplit(const gb_conv_params& p, cudaStream_t st);       // igemm_xsplit.cu: -1 = not applicable
int gb_conv_data_pair(const gb_conv_params& p, cudaStream_t st);  // igemm_pair.cu: -1 = not applicable
int gb_conv_data_cg2(const gb_conv_params& p, cudaStream_t st);   // igemm_cg2.cu: -1 = not applicable / switched off

extern "C" int gb_conv_data(const gb_conv_params* pp, void* stream) {
  const gb_conv_params& p = *pp;
  GB_CHECK(p.in.ptr && p.out.ptr && p.wpacked, "gb_conv_data: null pointer");
  GB_CHECK(p.in.C % 8 == 0 && p.out.C % 8 == 0, "gb_conv_data: channel counts must be multiples of 8 (%d, %d)", p.in.C,
           p.out.C);
  GB_CHECK(p.nclass >= 1 && p.nclass <= GB_MAX_CLASSES, "gb_conv_data: bad class count %d", p.nclass);
  GB_CHECK(p.ncols >= 1 && p.ncols <= p.out.C && p.npad % 16 == 0, "gb_conv_data: bad ncols/npad %d/%d", p.ncols,
           p.npad);
  GB_CHECK(p.in.N == p.out.N, "gb_conv_data: batch mismatch");
  GB_CHECK(!p.accumulate || p.out_fp32, "gb_conv_data: accumulate needs an fp32 output");
  GB_CHECK(view_max_offset(p.in) < (1ll << 31) && view_max_offset(p.out) < (1ll << 31),
           "gb_conv_data: tensor too large for 32-bit offsets");
  GB_CHECK(p.in.H < 32768 && p.in.W < 32768, "gb_conv_data: spatial extent too large");
  GB_CHECK(((uintptr_t)p.in.ptr & 15) == 0 && ((uintptr_t)p.out.ptr & 15) == 0 && ((uintptr_t)p.wpacked & 15) == 0,
           "gb_conv_data: pointers must be 16-byte aligned");
  int64_t max_mc = 0;
  ClassDivs divs;
  for (int c = 0; c < p.nclass; ++c) {
    GB_CHECK(p.cls[c].kpad % 64 == 0, "gb_conv_data: kpad must be a multiple of 64");
    GB_CHECK(p.cls[c].ntaps * p.in.C <= p.cls[c].kpad, "gb_conv_data: kpad too small");
    GB_CHECK(p.cls[c].tap_begin + p.cls[c].ntaps <= GB_MAX_TAPS, "gb_conv_data: too many taps");
    int q[3];
    gb_class_extents(p, c, q);
    int64_t mc = (int64_t)p.in.N * q[0] * q[1] * q[2];
    if (mc > max_mc) max_mc = mc;
    for (int d = 0; d < 3; ++d) divs.f[c][d] = gb_make_fastdiv((uint32_t)(q[d] > 0 ? q[d] : 1));
    GB_CHECK((int64_t)p.npad * p.cls[c].kpad < (1ll << 31), "gb_conv_data: packed weight matrix too large");
  }
  GB_CHECK(max_mc < (1ll << 31), "gb_conv_data: too many output positions");
  if (max_mc == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  {
    int r = gb_conv_data_halo(p, st);   // halo-reuse TMA kernel (dense tap windows, single class)
    if (r >= 0) return r;
    r = gb_conv_data_xsplit(p, st);       // 16 / 32 channels in, <= 32 out: the dx taps as columns of one MMA
    if (r >= 0) return r;
    r = gb_conv_data_halo_narrow(p, st);  // 16 / 32 input channels, one MMA per tap (what x-split does not take)
    if (r >= 0) return r;
    r = gb_conv_data_cg2(p, st);        // CTA-pair persistent kernel (opt-in: knob 16)
    if (r >= 0) return r;
    r = gb_conv_data_pair(p, st);       // two patches per CTA + halo reuse: wide stride-1 layers on full launches
    if (r >= 0) return r;
    r = gb_conv_data_tma(p, st);        // TMA-fed kernel for gathers with C % 64 == 0 (strided boxes for strides)
    if (r >= 0) return r;
  }
  GB_CHECK(p.in_c_valid == 0, "gb_conv_data: in_c_valid (pixel-window views) needs a TMA-fed kernel");
  // tile width: smallest BN covering the output channels, shrunk while the grid under-fills the 148 SMs
  int bn = 16;
  while (bn < p.ncols && bn < 256) bn *= 2;
  if (g_gb_knobs[1] > 0) {
    bn = g_gb_knobs[1];
  } else {
    const int64_t mt = (max_mc + BM - 1) / BM * p.nclass;
    while (bn > 64 && mt * gb_cdiv(p.ncols, bn) < 148) bn /= 2;
  }
  switch (bn) {
    case 16: return launch<16>(p, divs, max_mc, st);
    case 32: return launch<32>(p, divs, max_mc, st);
    case 64: return launch<64>(p, divs, max_mc, st);
    case 128: return launch<128>(p, divs, max_mc, st);
    case 256: return launch<256>(p, divs, max_mc, st);
  }
  GB_CHECK(false, "gb_conv_data: bad tile width %d", bn);
}
