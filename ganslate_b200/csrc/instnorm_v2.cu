// Second-generation fused InstanceNorm backward, opt-in (gb_debug_knob(22, v)): v = 1 -> 4 pixels (192 bytes) in
// flight per thread (3 with a residual gradient), v = 2 -> 2 pixels; two 256-thread blocks per SM (128 registers
// each, no spills) either way.  0 (default) keeps the
// first-generation kernels of instnorm_fast.cu, which are the ones measured in profiles/.  Knob 6 = 1 forces the
// two-launch form (no grid barrier), as for the first generation.
//
// The per-thread streaming body lives in instnorm_v2_core.h (compiled for the host as well: the CPU test-suite runs
// it thread by thread against torch, tests/test_in_bwd_v2_emul.py).  This file adds what only exists on the device:
// the block reduction of the partial sums, the atomics, the grid barrier of the single-launch form and the launch
// geometry (one co-resident wave, blocks of an image own equal pixel ranges).
#include "gb_common.cuh"
#include "instnorm_v2_core.h"

namespace {

using gbv2::Geom;
using gbv2::THREADS;

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

// partial sums (a, b) of the block's threads -> per-channel totals handed to `emit(channel, sum_a, sum_b)`;
// red: [slots][C][2] floats; ends with a block barrier so that `red` can be reused at once
template <typename Emit>
__device__ __forceinline__ void block_reduce(int C, const float (&a)[8], const float (&b)[8], float* red, Emit emit) {
  const int C8 = C >> 3;
  const int slots = THREADS / C8;
  const int cgp = threadIdx.x % C8, slot = threadIdx.x / C8;
  if (slot < slots) {
    float4* dst = reinterpret_cast<float4*>(red + ((size_t)slot * C + cgp * 8) * 2);
#pragma unroll
    for (int h = 0; h < 4; ++h) dst[h] = make_float4(a[2 * h], b[2 * h], a[2 * h + 1], b[2 * h + 1]);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += THREADS) {
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < slots; ++k) {
      const float2 t = *reinterpret_cast<const float2*>(red + ((size_t)k * C + ch) * 2);
      s1 += t.x;
      s2 += t.y;
    }
    emit(ch, s1, s2);
  }
  __syncthreads();
}

// PASS 0 / 1: the two passes as separate launches, PASS 2: both in one launch around a grid barrier.
// GEN: PReLU / residual before the activation / scaled output (instnorm_v2_core.h).
template <bool RES, int U, int MINB, int PASS, bool GEN>
__global__ void __launch_bounds__(THREADS, MINB)
in_bwd_v2_kernel(const __grid_constant__ gb_in_bwd_params p, const __grid_constant__ Geom g, float neg_slope) {
  gb_pdl_enter();
  extern __shared__ float red[];
  const int n = blockIdx.y;
  const int C = p.x.C;
  float acc1[8], acc2[8], acc3[8];
  if (PASS == 0 || PASS == 2) {
    gbv2::stream_pass<RES, U, 0, GEN>(p, g, neg_slope, threadIdx.x, blockIdx.x, n, acc1, acc2, acc3);
    float* bs = p.bstats + (int64_t)n * C * 2;
    block_reduce(C, acc1, acc2, red, [bs](int ch, float s1, float s2) {
      atomicAdd(bs + ch * 2 + 0, s1);
      atomicAdd(bs + ch * 2 + 1, s2);
    });
    if (GEN && p.dprelu != nullptr) {  // uniform over the grid
      float* dp = p.dprelu;
      block_reduce(C, acc3, acc3, red, [dp](int ch, float s1, float) { atomicAdd(dp + ch, s1); });
    }
  }
  if (PASS == 2) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(p.bstats + (int64_t)p.x.N * C * 2);
    grid_barrier(counter, (unsigned int)g.total_blocks);
  }
  if (PASS == 1 || PASS == 2) {
    gbv2::stream_pass<RES, U, 1, GEN>(p, g, neg_slope, threadIdx.x, blockIdx.x, n, acc1, acc2, acc3);
    if (p.dbias != nullptr) {  // uniform over the grid
      float* db = p.dbias;
      block_reduce(C, acc1, acc1, red, [db](int ch, float s1, float) { atomicAdd(db + ch, s1); });
    }
  }
}

bool row_addressable(const gb_view& v) { return v.D == 1 || (v.pad == 0 && v.sz == (int64_t)v.H * v.sy); }
bool small_offsets(const gb_view& v) {
  return ((int64_t)v.D * v.H + 2 * v.pad) * v.sy + (int64_t)(v.W + 2 * v.pad) * v.sx < (1ll << 31) &&
         (int64_t)v.D * v.H < (1 << 20) && v.W < (1 << 20);
}
bool aligned(const gb_view& v, int elem_bytes, int vec) {
  return ((uintptr_t)v.ptr % (elem_bytes * vec)) == 0 && v.sx % vec == 0 && v.sy % vec == 0 && v.sz % vec == 0 &&
         v.sn % vec == 0;
}
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <bool RES, int U, int MINB, bool GEN>
int launch(const gb_in_bwd_params& p, float neg_slope, cudaStream_t st) {
  const size_t smem = sizeof(float) * 2 * gbv2::slots_of(p.x.C) * p.x.C;
  static int occ = -1;  // co-resident blocks per SM of the single-launch kernel (per instantiation)
  static size_t occ_smem = 0;
  if (occ < 0 || occ_smem != smem) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, in_bwd_v2_kernel<RES, U, MINB, 2, GEN>, THREADS, smem) != cudaSuccess) o = 0;
    cudaGetLastError();
    occ = o;
    occ_smem = smem;
  }
  bool fits = false;
  const Geom g = gbv2::plan(p.x.N, p.x.D, p.x.H, p.x.W, p.x.C, num_sms() * (occ > 0 ? occ : MINB), &fits);
  const dim3 grid(g.nblocks, p.x.N);
  if (occ > 0 && fits && g_gb_knobs[6] == 2) {
    float ns = neg_slope;
    Geom gg = g;
    void* args[] = {(void*)&p, (void*)&gg, (void*)&ns};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)in_bwd_v2_kernel<RES, U, MINB, 2, GEN>, grid, dim3(THREADS), args,
                                                smem, st);
    if (e == cudaSuccess) {
      __atomic_fetch_add(&g_gb_launches, 1ull, __ATOMIC_RELAXED);
      return 0;
    }
    cudaGetLastError();  // cooperative launch not possible here: two launches
  }
  gb_klaunch(in_bwd_v2_kernel<RES, U, MINB, 0, GEN>, grid, THREADS, smem, st, p, g, neg_slope);
  GB_LAUNCH_CHECK();
  gb_klaunch(in_bwd_v2_kernel<RES, U, MINB, 1, GEN>, grid, THREADS, smem, st, p, g, neg_slope);
  GB_LAUNCH_CHECK();
  return 0;
}

template <int U, bool GEN>
__global__ void __launch_bounds__(THREADS, GEN ? 2 : 3)  // the general form keeps PReLU slopes + a residual vector per pixel
in_fwd_v2_kernel(const __grid_constant__ gb_in_fwd_params p, const __grid_constant__ Geom g, float neg_slope) {
  gb_pdl_enter();
  gbv2::fwd_pass<U, GEN>(p, g, neg_slope, threadIdx.x, blockIdx.x, blockIdx.y);
}

}  // namespace

// Forward, opt-in (gb_debug_knob(26, 1)): -1 = not covered, 0 = launched
int gb_in_fwd_fast_v2(const gb_in_fwd_params& p, cudaStream_t st) {
  if (g_gb_knobs[26] != 1) return -1;
  float ns;
  switch (p.act) {
    case GB_ACT_NONE: ns = 1.f; break;
    case GB_ACT_RELU: ns = 0.f; break;
    case GB_ACT_LEAKY: ns = p.act_slope; break;
    case GB_ACT_PRELU: ns = 0.f; break;
    default: return -1;
  }
  if (p.act == GB_ACT_PRELU && p.prelu == nullptr) return -1;
  const gb_view& x = p.x;
  const bool has_res = p.res.ptr != nullptr;
  const bool gen = p.act == GB_ACT_PRELU || (has_res && p.res_before_act != 0) || (p.out_scale != 0.f && p.out_scale != 1.f);
  if (x.C % 8 != 0 || x.C / 8 > THREADS || (int64_t)x.D * x.H * x.W >= (1ll << 31) || x.N > 65535) return -1;
  if (!aligned(x, 2, 8) || !aligned(p.y, 2, 8) || (has_res && !aligned(p.res, 2, 8))) return -1;
  if (p.stats != nullptr && (uintptr_t)p.stats % 16 != 0) return -1;
  if (!row_addressable(x) || !row_addressable(p.y) || (has_res && !row_addressable(p.res))) return -1;
  if (!small_offsets(x) || !small_offsets(p.y) || (has_res && !small_offsets(p.res))) return -1;
  if (p.y.pad > 0 && (p.y.D != 1 || p.y.H <= 2 * p.y.pad + 1 || p.y.W <= 2 * p.y.pad + 1)) return -1;
  bool fits = false;
  const Geom g = gbv2::plan(x.N, x.D, x.H, x.W, x.C, num_sms() * (gen ? 2 : 3), &fits);
  const dim3 grid(g.nblocks, x.N);
  if (gen) gb_klaunch(in_fwd_v2_kernel<4, true>, grid, THREADS, 0, st, p, g, ns);
  else gb_klaunch(in_fwd_v2_kernel<4, false>, grid, THREADS, 0, st, p, g, ns);
  GB_LAUNCH_CHECK();
  return 0;
}

// -1: not covered (caller falls back to the first-generation / general kernels), 0: launched, > 0: error
int gb_in_bwd_fast_v2(const gb_in_bwd_params& p, cudaStream_t st) {
  const int variant = g_gb_knobs[22];
  if (variant != 1 && variant != 2) return -1;
  float ns;
  switch (p.act) {
    case GB_ACT_NONE: ns = 1.f; break;
    case GB_ACT_RELU: ns = 0.f; break;
    case GB_ACT_LEAKY: ns = p.act_slope; break;
    case GB_ACT_PRELU: ns = 0.f; break;  // per-channel slopes from p.prelu
    default: return -1;
  }
  if (p.act == GB_ACT_PRELU && p.prelu == nullptr) return -1;
  if (p.stats == nullptr || p.bstats == nullptr) return -1;
  if (p.dy_a.ptr != nullptr || p.dy_b.ptr == nullptr) return -1;
  if (p.dx_fp32_acc) return -1;
  const bool rba = p.res_before_act != 0 && p.res.ptr != nullptr;
  // the general form: PReLU (+ its gradient), residual before the activation, scaled output
  const bool gen = p.act == GB_ACT_PRELU || rba || (p.out_scale != 0.f && p.out_scale != 1.f);
  if (!gen && p.dprelu != nullptr) return -1;
  const bool has_res = p.dy_sum.ptr != nullptr;
  if (has_res && !p.dy_sum_acc) return -1;
  const gb_view& x = p.x;
  if (x.C % 8 != 0 || x.C / 8 > THREADS || (int64_t)x.D * x.H * x.W >= (1ll << 31)) return -1;
  if (!aligned(x, 2, 8) || !aligned(p.dx, 2, 8) || !aligned(p.dy_b, 4, 4) || (has_res && !aligned(p.dy_sum, 4, 4))) return -1;
  if ((uintptr_t)p.stats % 16 != 0 || (uintptr_t)p.bstats % 16 != 0) return -1;
  if (!row_addressable(x) || !row_addressable(p.dx) || !row_addressable(p.dy_b) || (has_res && !row_addressable(p.dy_sum)))
    return -1;
  if (!small_offsets(x) || !small_offsets(p.dx) || !small_offsets(p.dy_b) || (has_res && !small_offsets(p.dy_sum))) return -1;
  if (p.dy_b.pad > 0 && (p.dy_b.D != 1 || p.dy_b.H <= 2 * p.dy_b.pad + 1 || p.dy_b.W <= 2 * p.dy_b.pad + 1)) return -1;
  if (rba && (!aligned(p.res, 2, 8) || !row_addressable(p.res) || !small_offsets(p.res))) return -1;
  if (gen) return has_res ? launch<true, 2, 2, true>(p, ns, st) : launch<false, 2, 2, true>(p, ns, st);
  // (the residual form holds two more fp32 vectors per pixel: one pixel less in flight keeps it free of spills)
  if (variant == 1) return has_res ? launch<true, 3, 2, false>(p, ns, st) : launch<false, 4, 2, false>(p, ns, st);
  return has_res ? launch<true, 2, 2, false>(p, ns, st) : launch<false, 2, 2, false>(p, ns, st);
}
