// Second-generation weight pack / weight-gradient unpack, opt-in (gb_debug_knob(28, 1)); element bodies in
// pack_v2_core.h (run on the CPU by tests/test_pack_v2_emul.py).
#include "gb_common.cuh"
#include "pack_v2_core.h"

namespace {

__global__ void __launch_bounds__(256) pack_multi_v2_kernel(const gb_pack_params* __restrict__ table) {
  gb_pdl_enter();
  const gb_pack_params& p = table[blockIdx.y];
  for (int cls = 0; cls < p.nclass; ++cls) {
    const uint32_t total8 = (uint32_t)p.rows_pad * ((uint32_t)p.kpad[cls] >> 3);
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dst) + p.w_offset[cls]);
    for (uint32_t i8 = blockIdx.x * blockDim.x + threadIdx.x; i8 < total8; i8 += gridDim.x * blockDim.x)
      dst[i8] = gbp2::pack8(p, cls, i8);
  }
}

__global__ void __launch_bounds__(256) unpack_multi_v2_kernel(const __grid_constant__ gb_unpack_batch b) {
  gb_pdl_enter();
  const gb_unpack_item& it = b.item[blockIdx.y];
  const uint32_t total = (uint32_t)it.rows * (uint32_t)it.chans * (uint32_t)it.ntaps;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) gbp2::unpack1(it, i);
}

}  // namespace

// -1: not covered (first generation takes the call), 0: launched.  The host cannot read the device-resident table, so
// the caller states what it guarantees: every class offset 16-byte aligned (w_offset % 8 == 0) and < 2^31 elements.
int gb_pack_weights_multi_v2(const gb_pack_params* table_dev, int count, int64_t max_elems, cudaStream_t st) {
  if (g_gb_knobs[28] == 2 || max_elems >= (1ll << 31)) return -1;
  int blocks = (int)((max_elems / 8 + 255) / 256);
  if (blocks > 592) blocks = 592;
  if (blocks < 1) blocks = 1;
  gb_klaunch(pack_multi_v2_kernel, dim3(blocks, count), 256, 0, st, table_dev);
  GB_LAUNCH_CHECK();
  return 0;
}

int gb_unpack_wgrad_multi_v2(const gb_unpack_batch* b, int64_t max_total, cudaStream_t st) {
  if (g_gb_knobs[28] == 2 || max_total >= (1ll << 31)) return -1;
  int blocks = (int)((max_total + 1023) / 1024);
  if (blocks > 592) blocks = 592;
  if (blocks < 1) blocks = 1;
  gb_klaunch(unpack_multi_v2_kernel, dim3(blocks, b->count), 256, 0, st, *b);
  GB_LAUNCH_CHECK();
  return 0;
}
