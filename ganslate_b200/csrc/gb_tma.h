// TMA helpers shared by the TMA-fed kernels: tensor-map construction (host) and bulk-tensor loads (device).
#pragma once
#include <cuda.h>
#include "gb_common.cuh"

// true when cuTensorMapEncodeTiled could be resolved from the driver (false on a CPU-only box)
bool gb_tma_available();
// 5-D map {C, W, H, D, N} of a channels-last bf16 view with box {64 ch, tw, th, 1, 1}, 128B swizzle, zero OOB fill.
// mul (optional, {z, y, x}): traversal strides -- the box delivers pixels c, c + mul, c + 2 mul ... (a strided
// convolution's gather as one TMA box).  c_valid (optional): channels that exist in memory (the map's channel extent;
// the 64-wide box zero-fills the rest).  Returns 0 on success (maps are cached by view + box + strides).
int gb_tma_activation_map(const gb_view& v, int tw, int th, CUtensorMap* out, const int* mul = nullptr, int c_valid = 0);

// 5-D map {C, W, H, D, N} of a channels-last OUTPUT view (bf16, or fp32 when `fp32` is set) for the TMA-store epilogue:
// box {128 bytes of channels (64 bf16 / 32 fp32), tw * mul_x, th * mul_y, mul_z, 1} with traversal strides `mul` (the
// output multiplier of a parity-class decomposed convolution: the box scatters to pixels c, c + mul, ...), 128B
// swizzle.  Stores clip at the tensor's extents, so tiles that overhang the image need no masking.
int gb_tma_store_map(const gb_view& v, int tw, int th, const int* mul, int fp32, CUtensorMap* out);

// 2-D fp32 map {cols, rows} (row pitch = cols * 4 bytes) with box {32 floats = 128 bytes, box_rows}, 128B swizzle: the
// weight-gradient workspace as the destination of bulk reduce-adds.
int gb_tma_f32_matrix_map(const void* ptr, int cols, int rows, int box_rows, CUtensorMap* out);

#ifdef __CUDACC__
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// smem (128B-swizzled tile, written by the generic proxy + fence.proxy.async) -> global, clipped at the tensor bounds
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
// same, global += smem (fp32 add performed at the L2)
__device__ __forceinline__ void tma_reduce_add_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3,
                                                  int c4) {
  asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// until the bulk stores of this thread have READ their shared-memory source (the CTA may then reuse / release it)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
#endif
