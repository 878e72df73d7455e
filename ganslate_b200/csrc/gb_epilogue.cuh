// Epilogue shared by the implicit-GEMM data kernels (gather, TMA, halo): TMEM accumulator tile ->
// (+bias, activation) -> bf16 or fp32 (optionally accumulated) global store, plus the optional per-(image,
// channel) InstanceNorm statistics of the bf16-rounded outputs.
#pragma once
#include "gb_common.cuh"

// Sum over the 32 lanes of a warp of CH per-lane values, for every column at once: after the butterfly lane L
// holds the total of column (L % CH) in v[0].  31 (CH=32) / 31 (CH=16) shuffles instead of 5 per column.
template <int CH>
__device__ __forceinline__ float gb_warp_colsum(float (&v)[CH], int lane) {
  static_assert(CH == 16 || CH == 32, "CH must be 16 or 32");
  if constexpr (CH == 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
  }
#pragma unroll
  for (int off = (CH == 32 ? 16 : 8); off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ float gb_round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Called by all 8 warps.  Warp w reads TMEM lanes (w & 3) * 32 .. +31 (its row = lane), the two warp halves
// (w >> 2) split the BN columns.  row_ok / ooff / row_n describe this thread's output row: valid, element offset of
// its pixel in p.out, image index.
// scratch != nullptr (kernels whose tile lies inside ONE image, BN >= 64): the statistics of the four row warps are
// summed in shared memory first (scratch: 4 * BN * 2 floats, free pipeline smem) and leave the CTA as one atomic per
// column and moment -- on the 256x256x64-channel layers 4096 CTAs otherwise queue 2 M atomics on 1024 addresses.
template <int BN>
__device__ __forceinline__ void gb_conv_epilogue(const gb_conv_params& p, uint32_t tmem_base, int warp, int lane,
                                                 bool have_acc, bool row_ok, int64_t ooff, int n0, const float* bias_s,
                                                 int row_n, float* scratch = nullptr, float* smem_stats = nullptr) {
  const int lg = warp & 3;
  const int half = warp >> 2;
  __nv_bfloat16* optr = reinterpret_cast<__nv_bfloat16*>(p.out.ptr);
  constexpr int CH = (BN >= 64) ? 32 : 16;  // columns per TMEM load
  constexpr int COLS_PER_HALF = (BN >= 64) ? BN / 2 : BN;
  const bool active = (BN >= 64) || half == 0;
  if (!active) return;
  const bool want_stats = p.stats != nullptr && !p.out_fp32;
  const bool cta_reduce = (BN >= 64) && want_stats && scratch != nullptr;
  // all valid rows of this warp belong to one image? (always true for the TMA tilings; the flat row tiling of
  // the gather kernel can straddle two images when an image is not a multiple of 32 pixels)
  bool uniform_n = true;
  int n_first = 0;
  if (want_stats) {
    const unsigned okmask = __ballot_sync(0xffffffffu, row_ok);
    if (okmask != 0u) {
      n_first = __shfl_sync(0xffffffffu, row_n, __ffs(okmask) - 1);
      uniform_n = __all_sync(0xffffffffu, !row_ok || row_n == n_first);
    }
  }
  const int cbeg = (BN >= 64) ? half * COLS_PER_HALF : 0;
#pragma unroll 1
  for (int c0 = cbeg; c0 < cbeg + COLS_PER_HALF; c0 += CH) {
    uint32_t acc[CH];
    if (have_acc) {
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0;
      if constexpr (CH == 32) tmem_ld32(taddr, acc); else tmem_ld16(taddr, acc);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i) acc[i] = 0u;
    }
    float sv[CH];  // bf16-rounded outputs of this row (0 for rows / columns that do not exist): statistics input
#pragma unroll
    for (int g = 0; g < CH / 8; ++g) {
      const int col = n0 + c0 + g * 8;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float t = __uint_as_float(acc[g * 8 + e]) + bias_s[c0 + g * 8 + e];
        if (p.act == GB_ACT_TANH) t = tanhf(t);
        else if (p.act == GB_ACT_LEAKY) t = t > 0.f ? t : t * p.act_slope;
        else if (p.act == GB_ACT_RELU) t = fmaxf(t, 0.f);
        v[e] = t;
      }
      const bool col_ok = row_ok && col < p.out.C;
      if (p.out_fp32) {
        if (col_ok) {
          float4* o32 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out.ptr) + ooff + col);
          float4 a = make_float4(v[0], v[1], v[2], v[3]), b = make_float4(v[4], v[5], v[6], v[7]);
          if (p.accumulate) {
            const float4 pa = o32[0], pb = o32[1];
            a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
            b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
          }
          o32[0] = a;
          o32[1] = b;
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
          f = unpack_bf16x2(o.x); sv[g * 8 + 0] = f.x; sv[g * 8 + 1] = f.y;
          f = unpack_bf16x2(o.y); sv[g * 8 + 2] = f.x; sv[g * 8 + 3] = f.y;
          f = unpack_bf16x2(o.z); sv[g * 8 + 4] = f.x; sv[g * 8 + 5] = f.y;
          f = unpack_bf16x2(o.w); sv[g * 8 + 6] = f.x; sv[g * 8 + 7] = f.y;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (!row_ok || col + e >= p.ncols) sv[g * 8 + e] = 0.f;
        }
      }
    }
    if (cta_reduce) {
      float sq[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) sq[i] = sv[i] * sv[i];
      const float s1 = gb_warp_colsum<CH>(sv, lane);
      const float s2 = gb_warp_colsum<CH>(sq, lane);
      if (lane < CH) {
        scratch[(lg * BN + c0 + lane) * 2 + 0] = s1;
        scratch[(lg * BN + c0 + lane) * 2 + 1] = s2;
      }
    } else if (want_stats) {
      if (uniform_n) {
        float sq[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) sq[i] = sv[i] * sv[i];
        const float s1 = gb_warp_colsum<CH>(sv, lane);
        const float s2 = gb_warp_colsum<CH>(sq, lane);
        const int col = n0 + c0 + (lane & (CH - 1));
        if (smem_stats != nullptr) {
          // narrow tiles, several accumulators per CTA (igemm_halo_narrow.cu): the sums of the row warps and patches
          // meet in shared memory, [BN][2] floats zeroed by the caller, who issues ONE global atomic per column and
          // moment -- a 32-column layer on 2 M voxels otherwise queues 17 M atomics on 64 addresses (measured 150 us)
          if (lane < CH) {
            atomicAdd(smem_stats + (c0 + lane) * 2, s1);
            atomicAdd(smem_stats + (c0 + lane) * 2 + 1, s2);
          }
        } else if (lane < CH && col < p.ncols) {
          float* dst = p.stats + ((int64_t)n_first * p.out.C + col) * 2;
          atomicAdd(dst, s1);
          atomicAdd(dst + 1, s2);
        }
      } else if (row_ok) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int col = n0 + c0 + i;
          if (col < p.ncols) {
            float* dst = p.stats + ((int64_t)row_n * p.out.C + col) * 2;
            atomicAdd(dst, sv[i]);
            atomicAdd(dst + 1, sv[i] * sv[i]);
          }
        }
      }
    }
  }
  if (cta_reduce) {  // uniform across the CTA: every warp of a BN >= 64 tile gets here
    __syncthreads();
    for (int i = warp * 32 + lane; i < BN; i += 256) {
      const int col = n0 + i;
      if (col < p.ncols) {
        const float a = scratch[i * 2] + scratch[(BN + i) * 2] + scratch[(2 * BN + i) * 2] + scratch[(3 * BN + i) * 2];
        const float b = scratch[i * 2 + 1] + scratch[(BN + i) * 2 + 1] + scratch[(2 * BN + i) * 2 + 1] +
                        scratch[(3 * BN + i) * 2 + 1];
        float* dst = p.stats + ((int64_t)row_n * p.out.C + col) * 2;
        atomicAdd(dst, a);
        atomicAdd(dst + 1, b);
      }
    }
    __syncthreads();  // scratch may be reused by a following call (second patch of the pair kernel)
  }
}
