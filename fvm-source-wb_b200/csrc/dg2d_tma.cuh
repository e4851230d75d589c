// TMA / mbarrier wrappers and the tensor-box width shared by the TMA-staged DG 2D stage kernel (dg2d_split.cuh).
// (Round 1's one-thread-per-element TMA kernel and its marching variant lived here; k_dg_stage_split replaced both: it keeps
//  their data path -- 3-D tensor boxes of 36 columns x 1 row x all planes, 16-byte aligned start, one mbarrier per slot --
//  and splits the element over four threads.)
#pragma once
#include <cuda.h>
#include <cstdint>

namespace wb { namespace dg {

constexpr int DGT_W = 36;                     // box columns: 2 (alignment) + 32 + 2

namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
      : "memory");
}
}  // namespace tma

}}  // namespace wb::dg
