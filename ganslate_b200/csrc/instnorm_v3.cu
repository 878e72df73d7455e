// On-chip fused InstanceNorm backward, opt-in (gb_debug_knob(24, 1); 24 = 2: half-size stash, two CTAs per SM): a thread-block cluster of K <= 8 CTAs owns
// (image, 32 channels), streams its pixels ONCE into shared memory, exchanges the partial sums through distributed
// shared memory and writes dx from the shared-memory copy.  See instnorm_v3_core.h for the design and the per-thread
// body (which the CPU test-suite runs on the host, tests/test_in_bwd_v2_emul.py); this file holds the device-only part:
// warp / block reduction, the cluster exchange, the launch.
//
// Shared memory of a CTA: stash [steps][256] x (float4, float4, uint4) <= 192 KB, then
//   part_w [8 warps][32 channels][2] floats   per-warp partial sums
//   xchg   [32][2] floats                     this CTA's partial sums, read by the cluster's other CTAs
//   tot    [32][2] floats                     cluster totals
#include "gb_common.cuh"
#include "instnorm_v3_core.h"

namespace {

using gbv3::CG;
using gbv3::Geom;
using gbv3::THREADS;
using gbv3::TPP;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float ld_peer_f32(const float* own_smem, uint32_t rank) {
  uint32_t a = smem_u32(own_smem), r;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(r) : "memory");
  return v;
}

// lanes with equal (lane % TPP) own the same 8 channels: sum over the 8 pixel slots of the warp, result in lanes 0..3
__device__ __forceinline__ void warp_sum_slots(float (&v)[8]) {
#pragma unroll
  for (int off = TPP; off < 32; off <<= 1)
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += __shfl_xor_sync(0xffffffffu, v[e], off);
}

template <bool RES>
__global__ void __launch_bounds__(THREADS, 1)
in_bwd_v3_kernel(const __grid_constant__ gb_in_bwd_params p, const __grid_constant__ Geom g, float neg_slope) {
  gb_pdl_enter();
  extern __shared__ __align__(16) uint8_t smem[];
  float4* st_g0 = reinterpret_cast<float4*>(smem);
  float4* st_g1 = st_g0 + (size_t)g.steps * THREADS;
  uint4* st_x = reinterpret_cast<uint4*>(st_g1 + (size_t)g.steps * THREADS);
  float* part_w = reinterpret_cast<float*>(st_x + (size_t)g.steps * THREADS);  // [8][CG][2]
  float* xchg = part_w + 8 * CG * 2;                                           // [CG][2]
  float* tot = xchg + CG * 2;                                                  // [CG][2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_rank();
  const int cgi = blockIdx.y, n = blockIdx.z;

  float s1[8], s2[8];
  gbv3::load_pass<RES, 4>(p, g, neg_slope, tid, rank, cgi, n, st_g0, st_g1, st_x, s1, s2);
  warp_sum_slots(s1);
  warp_sum_slots(s2);
  if (lane < TPP) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      part_w[(warp * CG + lane * 8 + e) * 2 + 0] = s1[e];
      part_w[(warp * CG + lane * 8 + e) * 2 + 1] = s2[e];
    }
  }
  __syncthreads();
  if (tid < CG * 2) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) a += part_w[w * CG * 2 + tid];
    xchg[tid] = a;
  }
  // every CTA's partial sums are in its xchg: make them visible to the cluster, then read all K of them in rank order
  // (every CTA adds in the same order: identical totals in all of them)
  cluster_arrive();
  cluster_wait();
  if (tid < CG * 2) {
    float a = 0.f;
    for (int r = 0; r < g.K; ++r) a += ld_peer_f32(xchg + tid, (uint32_t)r);
    tot[tid] = a;
  }
  cluster_arrive();  // this CTA has read its peers' xchg (matched by the wait before exit)
  __syncthreads();
  float t1[8], t2[8], db[8];
  {
    const int quad = tid % TPP;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      t1[e] = tot[(quad * 8 + e) * 2 + 0];
      t2[e] = tot[(quad * 8 + e) * 2 + 1];
    }
  }
  gbv3::apply_pass(p, g, tid, rank, cgi, n, t1, t2, st_g0, st_g1, st_x, db);
  if (p.dbias != nullptr) {  // uniform
    warp_sum_slots(db);
    if (lane < TPP) {
#pragma unroll
      for (int e = 0; e < 8; ++e) part_w[warp * CG + lane * 8 + e] = db[e];  // part_w's pass-1 contents were consumed before the cluster barrier
    }
    __syncthreads();
    if (tid < CG) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) a += part_w[w * CG + tid];
      atomicAdd(p.dbias + cgi * CG + tid, a);
    }
  }
  cluster_wait();  // no CTA leaves while a peer may still read its xchg
}

bool row_addressable(const gb_view& v) { return v.D == 1 || (v.pad == 0 && v.sz == (int64_t)v.H * v.sy); }
bool small_offsets(const gb_view& v) {
  return ((int64_t)v.D * v.H + 2 * v.pad) * v.sy + (int64_t)(v.W + 2 * v.pad) * v.sx < (1ll << 31);
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

template <bool RES>
int launch(const gb_in_bwd_params& p, const Geom& g, float neg_slope, cudaStream_t st) {
  const size_t smem = (size_t)g.steps * THREADS * 48 + sizeof(float) * (8 * CG * 2 + CG * 2 + CG * 2);
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    GB_CUDA(cudaFuncSetAttribute(in_bwd_v3_kernel<RES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 gbv3::MAX_STEPS * THREADS * 48 + 4096));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.K, p.x.C / CG, p.x.N);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = g.K;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.numAttrs = 1;
  if (g_gb_knobs[20] != 0) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cfg.attrs = attr;
  if (cudaLaunchKernelEx(&cfg, in_bwd_v3_kernel<RES>, p, g, neg_slope) != cudaSuccess) {
    // e.g. no GPC with K free SMs for a 199 KB CTA each (MIG slices, other work resident): the two-pass kernels
    // take the call; knob 25 (launches served here) lets the tests see that this happened
    cudaGetLastError();
    return -1;
  }
  __atomic_fetch_add(&g_gb_launches, 1ull, __ATOMIC_RELAXED);
  return 0;
}

}  // namespace

// -1: not covered (caller falls back to the two-pass kernels), 0: launched, > 0: error
int gb_in_bwd_onchip(const gb_in_bwd_params& p, cudaStream_t st) {
  if (g_gb_knobs[24] != 1 && g_gb_knobs[24] != 2) return -1;
  float ns;
  switch (p.act) {
    case GB_ACT_NONE: ns = 1.f; break;
    case GB_ACT_RELU: ns = 0.f; break;
    case GB_ACT_LEAKY: ns = p.act_slope; break;
    default: return -1;
  }
  if (p.stats == nullptr) return -1;
  if (p.dy_a.ptr != nullptr || p.dy_b.ptr == nullptr) return -1;
  if (p.res_before_act || p.dx_fp32_acc || p.dprelu != nullptr) return -1;
  if (p.out_scale != 0.f && p.out_scale != 1.f) return -1;
  const bool has_res = p.dy_sum.ptr != nullptr;
  if (has_res && !p.dy_sum_acc) return -1;
  const gb_view& x = p.x;
  if (p.x.N > 65535 || x.C / CG > 65535) return -1;
  if (!aligned(x, 2, 8) || !aligned(p.dx, 2, 8) || !aligned(p.dy_b, 4, 4) || (has_res && !aligned(p.dy_sum, 4, 4))) return -1;
  if ((uintptr_t)p.stats % 16 != 0) return -1;
  if (!row_addressable(x) || !row_addressable(p.dx) || !row_addressable(p.dy_b) || (has_res && !row_addressable(p.dy_sum)))
    return -1;
  if (!small_offsets(x) || !small_offsets(p.dx) || !small_offsets(p.dy_b) || (has_res && !small_offsets(p.dy_sum))) return -1;
  if (p.dy_b.pad > 0 && (p.dy_b.D != 1 || p.dy_b.H <= 2 * p.dy_b.pad + 1 || p.dy_b.W <= 2 * p.dy_b.pad + 1)) return -1;
  Geom g;
  if (!gbv3::plan(x.N, x.D, x.H, x.W, x.C, num_sms(), g_gb_knobs[24], &g)) return -1;
  return has_res ? launch<true>(p, g, ns, st) : launch<false>(p, g, ns, st);
}
