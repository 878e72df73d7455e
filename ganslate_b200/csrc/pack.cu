// Weight packing (fp32 PyTorch layout -> bf16 class matrices), weight-gradient unpacking and the
// per-channel column sum used for bias gradients.  All HBM-bound, vectorised where the layout allows.
#include "gb_common.cuh"

namespace {

__device__ __forceinline__ void pack_class(const gb_pack_params& p, int cls, int64_t first, int64_t step) {
  const int kpad = p.kpad[cls];
  const int64_t total = (int64_t)p.rows_pad * kpad;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.dst) + p.w_offset[cls];
  const int ntaps = p.ntaps[cls], tb = p.tap_begin[cls];
  for (int64_t i = first; i < total; i += step) {
    const int n = (int)(i / kpad);
    const int k = (int)(i - (int64_t)n * kpad);
    const int tl = k / p.chans_pad;
    const int c = k - tl * p.chans_pad;
    float v = 0.f;
    if (n < p.rows && tl < ntaps && c < p.chans) {
      const int t = p.tap_id[tb + tl];  // < 0: a padding tap of a pixel-window layout (stays zero)
      if (t >= 0) v = p.src[(int64_t)n * p.sn + (int64_t)c * p.sc + (int64_t)t * p.st];
    }
    dst[i] = __float2bfloat16_rn(v);
  }
}

// one launch for every convolution of a network: the table lives in device memory (built once per network, the
// parameter / packed-buffer addresses never change), blockIdx.y selects the entry
__global__ void pack_multi_kernel(const gb_pack_params* __restrict__ table) {
  gb_pdl_enter();
  const gb_pack_params& p = table[blockIdx.y];
  for (int cls = 0; cls < p.nclass; ++cls)
    pack_class(p, cls, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

__global__ void unpack_multi_kernel(const __grid_constant__ gb_unpack_batch b) {
  gb_pdl_enter();
  const gb_unpack_item& it = b.item[blockIdx.y];
  const int64_t total = (int64_t)it.rows * it.chans * it.ntaps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % it.ntaps);
    const int64_t rc = i / it.ntaps;
    const int c = (int)(rc % it.chans);
    const int r = (int)(rc / it.chans);
    const float v = it.dw[(int64_t)r * it.kpad + (int64_t)t * it.chans_pad + c];
    float* d = it.dst + (int64_t)r * it.dsr + (int64_t)c * it.dsc + (int64_t)t * it.dst_t;
    *d = it.accumulate ? *d + v : v;
  }
}

__global__ void pack_kernel(const __grid_constant__ gb_pack_params p) {
  gb_pdl_enter();
  pack_class(p, blockIdx.y, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

__global__ void unpack_kernel(const float* __restrict__ dw, float* __restrict__ dst, int64_t dsr, int64_t dsc,
                              int64_t dst_t, int rows, int chans, int chans_pad, int ntaps, int kpad) {
  gb_pdl_enter();
  const int64_t total = (int64_t)rows * chans * ntaps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // iterate in destination order (PyTorch layout is usually (r, c, t) contiguous) for coalesced writes
    const int t = (int)(i % ntaps);
    const int64_t rc = i / ntaps;
    const int c = (int)(rc % chans);
    const int r = (int)(rc / chans);
    dst[(int64_t)r * dsr + (int64_t)c * dsc + (int64_t)t * dst_t] = dw[(int64_t)r * kpad + (int64_t)t * chans_pad + c];
  }
}

// per-channel sums of a channels-last bf16 view. Block = 256 threads; thread = (pixel slot, 8-channel group).
__global__ void colsum_kernel(gb_view x, float* __restrict__ out, int pix_per_block) {
  gb_pdl_enter();
  extern __shared__ float red[];  // [slots][C]
  const int C8 = x.C >> 3;
  const int slots = blockDim.x / C8;
  const int cg = threadIdx.x % C8;
  const int slot = threadIdx.x / C8;
  const int64_t P = (int64_t)x.N * x.D * x.H * x.W;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p1 = min(P, p0 + pix_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const __nv_bfloat16* ptr = reinterpret_cast<const __nv_bfloat16*>(x.ptr);
  if (slot < slots) {
    for (int64_t pix = p0 + slot; pix < p1; pix += slots) {
      int64_t m = pix;
      const int xx = (int)(m % x.W); m /= x.W;
      const int yy = (int)(m % x.H); m /= x.H;
      const int zz = (int)(m % x.D);
      const int nn = (int)(m / x.D);
      const uint4 v = *reinterpret_cast<const uint4*>(ptr + nn * x.sn + zz * x.sz + yy * x.sy + xx * x.sx + cg * 8);
      float2 f;
      f = unpack_bf16x2(v.x); acc[0] += f.x; acc[1] += f.y;
      f = unpack_bf16x2(v.y); acc[2] += f.x; acc[3] += f.y;
      f = unpack_bf16x2(v.z); acc[4] += f.x; acc[5] += f.y;
      f = unpack_bf16x2(v.w); acc[6] += f.x; acc[7] += f.y;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[slot * x.C + cg * 8 + e] = acc[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < x.C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < slots; ++k) s += red[k * x.C + c];
    atomicAdd(out + c, s);
  }
}

}  // namespace

int gb_pack_weights_multi_v2(const gb_pack_params* table_dev, int count, int64_t max_elems, cudaStream_t st);  // pack_v2.cu
int gb_unpack_wgrad_multi_v2(const gb_unpack_batch* b, int64_t max_total, cudaStream_t st);                       // (knob 28)

extern "C" int gb_pack_weights(const gb_pack_params* pp, void* stream) {
  const gb_pack_params& p = *pp;
  GB_CHECK(p.src && p.dst, "gb_pack_weights: null pointer");
  GB_CHECK(p.nclass >= 1 && p.nclass <= GB_MAX_CLASSES, "gb_pack_weights: bad class count");
  GB_CHECK(p.chans_pad % 8 == 0 && p.rows_pad % 16 == 0, "gb_pack_weights: bad padding");
  int64_t mx = 0;
  for (int c = 0; c < p.nclass; ++c) {
    GB_CHECK(p.kpad[c] % 64 == 0 && p.ntaps[c] * p.chans_pad <= p.kpad[c], "gb_pack_weights: bad kpad");
    mx = mx > (int64_t)p.rows_pad * p.kpad[c] ? mx : (int64_t)p.rows_pad * p.kpad[c];
  }
  int blocks = (int)((mx + 255) / 256);
  if (blocks > 2048) blocks = 2048;
  if (blocks < 1) blocks = 1;
  gb_klaunch(pack_kernel, dim3(blocks, p.nclass), 256, 0, (cudaStream_t)stream, p);
  GB_LAUNCH_CHECK();
  return 0;
}

extern "C" int gb_pack_weights_multi(const gb_pack_params* table_dev, int count, int64_t max_elems, void* stream) {
  GB_CHECK(table_dev && count >= 1 && count <= 65535, "gb_pack_weights_multi: bad table (%d entries)", count);
  if (g_gb_knobs[28] != 2) {   // second-generation pack / unpack by default (r02a: +0.9 %, r02n: +2.4 %); knob 28 = 2: first generation
    const int r = gb_pack_weights_multi_v2(table_dev, count, max_elems, (cudaStream_t)stream);
    if (r >= 0) return r;
  }
  int blocks = (int)((max_elems + 1023) / 1024);
  if (blocks > 592) blocks = 592;
  if (blocks < 1) blocks = 1;
  gb_klaunch(pack_multi_kernel, dim3(blocks, count), 256, 0, (cudaStream_t)stream, table_dev);
  GB_LAUNCH_CHECK();
  return 0;
}

extern "C" int gb_unpack_wgrad_multi(const gb_unpack_batch* b, void* stream) {
  GB_CHECK(b && b->count >= 1 && b->count <= GB_UNPACK_BATCH, "gb_unpack_wgrad_multi: bad item count");
  int64_t mx = 0;
  for (int i = 0; i < b->count; ++i) {
    const gb_unpack_item& it = b->item[i];
    GB_CHECK(it.dw && it.dst, "gb_unpack_wgrad_multi: null pointer in item %d", i);
    const int64_t total = (int64_t)it.rows * it.chans * it.ntaps;
    mx = total > mx ? total : mx;
  }
  if (g_gb_knobs[28] != 2) {   // second-generation pack / unpack by default (r02a: +0.9 %, r02n: +2.4 %); knob 28 = 2: first generation
    const int r = gb_unpack_wgrad_multi_v2(b, mx, (cudaStream_t)stream);
    if (r >= 0) return r;
  }
  int blocks = (int)((mx + 1023) / 1024);
  if (blocks > 592) blocks = 592;
  if (blocks < 1) blocks = 1;
  gb_klaunch(unpack_multi_kernel, dim3(blocks, b->count), 256, 0, (cudaStream_t)stream, *b);
  GB_LAUNCH_CHECK();
  return 0;
}

extern "C" int gb_unpack_wgrad(const float* dw, float* dst, int64_t dsr, int64_t dsc, int64_t dst_t, int rows,
                               int chans, int chans_pad, int ntaps, int kpad, void* stream) {
  GB_CHECK(dw && dst, "gb_unpack_wgrad: null pointer");
  const int64_t total = (int64_t)rows * chans * ntaps;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  if (blocks < 1) blocks = 1;
  gb_klaunch(unpack_kernel, blocks, 256, 0, (cudaStream_t)stream, dw, dst, dsr, dsc, dst_t, rows, chans, chans_pad, ntaps, kpad);
  GB_LAUNCH_CHECK();
  return 0;
}

extern "C" int gb_colsum(const gb_view* x, float* out, void* stream) {
  GB_CHECK(x && x->ptr && out, "gb_colsum: null pointer");
  GB_CHECK(x->C % 8 == 0 && x->C <= 2048, "gb_colsum: bad channel count %d", x->C);
  cudaStream_t st = (cudaStream_t)stream;
  GB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * x->C, st));
  const int64_t P = (int64_t)x->N * x->D * x->H * x->W;
  const int C8 = x->C / 8;
  int threads = 256;
  if (C8 > threads) threads = ((C8 + 31) / 32) * 32;
  const int slots = threads / C8;
  int64_t blocks = (P + 63) / 64;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const int ppb = (int)((P + blocks - 1) / blocks);
  blocks = (P + ppb - 1) / ppb;
  gb_klaunch(colsum_kernel, (int)blocks, threads, sizeof(float) * slots * x->C, st, *x, out, ppb);
  GB_LAUNCH_CHECK();
  return 0;
}
