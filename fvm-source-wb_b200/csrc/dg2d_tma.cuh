// TMA-staged variant of the fused DG 2D RK-stage kernel.
//
// k_dg_stage_fast waits on memory (ncu: long_scoreboard 2.4 warps/issue at 8 warps/SM): every face first loads the 36 modes
// of the neighbour across it with per-plane address arithmetic, and with 255 registers there is no room to issue those
// loads early.  Here a block is ONE warp = 32 consecutive elements of a row, and its lane 0 asks the TMA unit for the rows
// it needs as 3-D tensor boxes (36 columns x 1 row x all 4*M*M planes, starting two columns left of the block so that the
// byte offset is 16-byte aligned):
//   region R1 <- own row (own modes + left/right neighbours), region R2 <- row below; when the x faces are done R1 is
//   re-armed with the row above.  Completion on one mbarrier per load; all reads are LDS with compile-time offsets.
// The x neighbours that wrap around the domain (ic = 0, nx-1) are not contiguous with the row: those two lanes take the
// global-memory path for that face.  Needs nx % 32 == 0; other grids use k_dg_stage_fast.
// Included by dg2d.cu after dg2d_fast.cuh; the mbarrier / TMA wrappers come from fv2d_tma.cuh's namespace.
#pragma once
#include <cuda.h>
#include <cstdint>

namespace wb { namespace dg {

#ifndef DGT_STAGE_RK
#define DGT_STAGE_RK 0      /* measured on B200 (4096^2, order 3): 2.84e9 without, 2.19e9 with -- see stage_operand below */
#endif
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

template <int M>
struct SmemSrc {
  static constexpr int NP = 4 * M * M;
  static constexpr int REGION_B = NP * DGT_W * 8;
  const double* __restrict__ in;
  const DgGrid& g;
  const CUtensorMap* map;
  const double* R1;            // own row, later the row above
  const double* R2;            // row below
  uint32_t r1, bars;           // shared-space addresses
  int lane, ic0, jt;
  size_t e;
  size_t eL, eR;               // wrapped x neighbours (used by the two edge lanes of a row only)
  bool left_edge, right_edge;

  __device__ __forceinline__ void from_smem(const double* R, int v, int col, double (&d)[M][M]) const {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int i = 0; i < M; ++i) d[i][j] = R[(v * M * M + j * M + i) * DGT_W + col];
  }
  __device__ __forceinline__ void own(int v, double (&d)[M][M]) const {
    if (v == 0) tma::mbar_wait(bars, 0);
    from_smem(R1, v, lane + 2, d);
  }
  template <int FACE>
  __device__ __forceinline__ void nb(int v, double (&d)[M][M]) const {
    if (FACE == 0) { if (left_edge) load_var<M>(in, g, v, eL, d); else from_smem(R1, v, lane + 1, d); }
    if (FACE == 1) { if (right_edge) load_var<M>(in, g, v, eR, d); else from_smem(R1, v, lane + 3, d); }
    if (FACE == 2) { if (v == 0) tma::mbar_wait(bars + 8, 0); from_smem(R2, v, lane + 2, d); }
    if (FACE == 3) { if (v == 0) tma::mbar_wait(bars + 16, 0); from_smem(R1, v, lane + 2, d); }
  }
  // own modes are in registers and both x faces are done: R1 is free for the row above
  __device__ __forceinline__ void x_faces_done() const {
    __syncwarp();
    if (lane == 0) {
      tma::mbar_expect_tx(bars + 16, REGION_B);
      tma::load_3d(r1, map, ic0 - 2, jt, 0, bars + 16);
    }
  }
  // (staging the RK operands A0 / in through TMA as well was measured: 1.99e9 instead of 2.59e9 element-stages/s)
#if DGT_STAGE_RK
  // EXPERIMENT, off by default (slower, like the TMA-staged variant tried earlier: 1.99e9).
  // The RK operands are needed at the very end of the block's life and their first use waits a full DRAM round trip
  // (ncu: 13 % of the stall samples).  Each lane copies its own 36 values with cp.async into a region that has become
  // free -- A0 into R2 once the bottom face is done, A1 into R1 once the top face is done -- and reads them back with LDS.
  // A lane only reads what it copied itself, so cp.async.wait_group is all the synchronisation needed after the copy.
  __device__ __forceinline__ void stage_operand(const double* A, uint32_t region) const {
    const uint32_t dst = region + (lane + 2) * 8;
#pragma unroll
    for (int k = 0; k < NP; ++k)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + k * DGT_W * 8), "l"(A + (size_t)k * g.ne + e));
    asm volatile("cp.async.commit_group;");
  }
  __device__ __forceinline__ void bottom_face_done(const StageCoef& C) const {
    __syncwarp();                                  // every lane has read the row below
    stage_operand(C.A0, r1 + REGION_B);
  }
  __device__ __forceinline__ void top_face_done(const StageCoef& C) const {
    __syncwarp();                                  // every lane has read the row above
    if (C.na >= 2) stage_operand(C.A1, r1);
  }
  __device__ __forceinline__ double rk_a0(const StageCoef& C, int v, int m) const {
    if (v == 0 && m == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    return R2[(v * M * M + m) * DGT_W + lane + 2];
  }
  __device__ __forceinline__ double rk_a1(const StageCoef& C, int v, int m) const { return R1[(v * M * M + m) * DGT_W + lane + 2]; }
#else
  __device__ __forceinline__ void bottom_face_done(const StageCoef&) const {}
  __device__ __forceinline__ void top_face_done(const StageCoef&) const {}
  __device__ __forceinline__ double rk_a0(const StageCoef& C, int v, int m) const { return PL(C.A0, g, v, m)[e]; }
  __device__ __forceinline__ double rk_a1(const StageCoef& C, int v, int m) const { return PL(C.A1, g, v, m)[e]; }
#endif
  __device__ __forceinline__ double rk_in(int v, int m) const { return PL(in, g, v, m)[e]; }
};

#ifndef DGT_MINB
#define DGT_MINB 1
#endif
template <int M, bool ANYFLUX>
__global__ void __launch_bounds__(32, DGT_MINB) k_dg_stage_tma(const __grid_constant__ CUtensorMap m_in, const double* __restrict__ in, StageCoef C,
                                                     double* __restrict__ out, const double* __restrict__ gx,
                                                     const double* __restrict__ gy, const unsigned char* __restrict__ fz, DgGrid g,
                                                     DgPhys P, FastBasis B, const DgCtrl* __restrict__ ctrl, int apply_onp) {
  extern __shared__ __align__(128) unsigned char dgt_smem[];
  if (ctrl->skip) { dg_stage_pass_through<M>(in, out, g, (size_t)blockIdx.x * 32 + threadIdx.x); return; }
  constexpr int REGION_B = SmemSrc<M>::REGION_B;
  const int lane = threadIdx.x;
  const size_t e0 = (size_t)blockIdx.x * 32;          // nx % 32 == 0: the 32 elements of a block lie in one row
  const int jc = (int)(e0 / g.nx), ic0 = (int)(e0 % g.nx), ic = ic0 + lane;
  const size_t e = e0 + lane;
  const int jb = y_nb(g, P.bc, jc - 1), jt = y_nb(g, P.bc, jc + 1);
  const uint32_t r1 = tma::smem_u32(dgt_smem), r2 = r1 + REGION_B, bars = r2 + REGION_B;
  if (lane == 0) {
    tma::mbar_init(bars, 1); tma::mbar_init(bars + 8, 1); tma::mbar_init(bars + 16, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma::mbar_expect_tx(bars, REGION_B);
    tma::load_3d(r1, &m_in, ic0 - 2, jc, 0, bars);
    tma::mbar_expect_tx(bars + 8, REGION_B);
    tma::load_3d(r2, &m_in, ic0 - 2, jb, 0, bars + 8);
  }
  __syncwarp();
  SmemSrc<M> src{in, g, &m_in, reinterpret_cast<const double*>(dgt_smem), reinterpret_cast<const double*>(dgt_smem + REGION_B),
                 r1, bars, lane, ic0, jt, e,
                 (size_t)jc * g.nx + bc_index(P.bc, ic - 1, g.nyg), (size_t)jc * g.nx + bc_index(P.bc, ic + 1, g.nyg),
                 ic == 0, ic == g.nx - 1};
  dg_stage_body<M, ANYFLUX>(src, in, C, out, gx, gy, fz, g, P, B, ctrl, apply_onp, e);
}

template <int M>
constexpr int dg_tma_smem_bytes() { return 2 * SmemSrc<M>::REGION_B + 128; }

}}  // namespace wb::dg
