// Marching variant of the fused DG 2D RK-stage kernel: every face flux is evaluated ONCE.
// OPT-IN (WB_DG2D_MARCH=1): measured on B200 it executes 18 % fewer instructions than k_dg_stage_tma and is still
// ~7 % slower (2.45e9 vs 2.62e9 element-stages/s at 4096^2, order 3) -- the stage is bound by latency at 8 warps/SM
// (255 registers), not by instruction issue, and a warp that walks a strip serialises its rows.  Kept because it is
// bit-identical to the two-sided kernels and is the starting point for a lower-register design.
//
// k_dg_stage_fast / k_dg_stage_tma evaluate each face from both of its elements (4 faces per element, no inter-thread
// traffic); the face terms are ~60 % of the stage's FP64 instructions and the kernel is bound by FP64 issue, not by HBM.
// Here a block is one warp that owns a strip of 31 element columns and `rows` element rows and walks up the strip:
//   * lane l evaluates the LEFT x face of element (ic0+l, j); the right face of that element is the left face of lane l+1
//     (__shfl_down).  Lane 31 (and the lane of the virtual element behind the last column) only feeds its left neighbour:
//     1/32 of the lanes idle in the volume part instead of one redundant face per element;
//   * the TOP y face flux is kept in registers and becomes the BOTTOM face flux of the next row; only the first row of a
//     strip evaluates its bottom face (1/rows redundancy).
// => 2 + 1/32 + 1/rows face evaluations per element instead of 4, bit-identical numbers (same traces, same LLF call with
//    the low side first, same accumulation order left, right, bottom, top).
// Data path: two shared-memory row slots fed by TMA (3-D boxes of 36 columns x 1 row x all planes, 16-byte aligned start).
// A row is loaded once per strip: slot s holds the own row, slot 1-s the row above; at the end of the iteration slot s is
// re-armed with row j+2 (needed half an iteration later) and the roles swap.  The own row stays staged until then so that
// the RK operands that alias the stage input are read from it instead of from global memory.  The x neighbours that wrap around the domain come from
// global memory.  Needs nx even (16-byte row pitch) and nx >= DGT_W; other grids use k_dg_stage_fast.
// Included by dg2d.cu after dg2d_fast.cuh and dg2d_tma.cuh (DGT_W, tma:: wrappers).
#pragma once

namespace wb { namespace dg {

constexpr int DGM_COLS = 31;                  // element columns owned by a warp

// RK operands of the element.  An operand that IS the stage input (A0 in stage 1, A1 in stages 2-4, `in` of the second
// result) is read back from the staged own row in shared memory: by the time the RK combination runs, the row has
// travelled ~2 row iterations through L2 and a global re-read would go to DRAM again (ncu: +43 % DRAM reads, 24 % of the
// stall samples on the first use of A0).  The others come from global memory.
template <int M>
struct MarchRk {
  const double* __restrict__ in;
  const DgGrid& g;
  size_t e;
  const double* own;                           // own row slot, this lane's column
  bool a0_in, a1_in;
  __device__ __forceinline__ double staged(int v, int m) const { return own[(v * M * M + m) * DGT_W]; }
  __device__ __forceinline__ double rk_a0(const StageCoef& C, int v, int m) const { return a0_in ? staged(v, m) : PL(C.A0, g, v, m)[e]; }
  __device__ __forceinline__ double rk_a1(const StageCoef& C, int v, int m) const { return a1_in ? staged(v, m) : PL(C.A1, g, v, m)[e]; }
  __device__ __forceinline__ double rk_in(int v, int m) const { return staged(v, m); }
};

template <int M>
__device__ __forceinline__ void march_from_smem(const double* R, int v, int col, double (&d)[M][M]) {
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int i = 0; i < M; ++i) d[i][j] = R[(v * M * M + j * M + i) * DGT_W + col];
}

template <int M, bool ANYFLUX>
__global__ void __launch_bounds__(32) k_dg_stage_march(const __grid_constant__ CUtensorMap m_in, const double* __restrict__ in,
                                                       StageCoef C, double* __restrict__ out, const double* __restrict__ gx,
                                                       const double* __restrict__ gy, const unsigned char* __restrict__ fz,
                                                       DgGrid g, DgPhys P, FastBasis B, const DgCtrl* __restrict__ ctrl,
                                                       int apply_onp, int rows) {
  extern __shared__ __align__(128) unsigned char dgm_smem[];
  constexpr int REGION_B = 4 * M * M * DGT_W * 8;
  const int lane = threadIdx.x;
  const int ic0 = blockIdx.x * DGM_COLS, ic = ic0 + lane;
  const int j0 = blockIdx.y * rows, j1 = min(j0 + rows, g.ny);
  const bool owner = lane < DGM_COLS && ic < g.nx;          // stores an element
  if (ctrl->skip) {
    if (owner)
      for (int jc = j0; jc < j1; ++jc) dg_stage_pass_through<M>(in, out, g, (size_t)jc * g.nx + ic);
    return;
  }
  const bool face_lane = ic <= g.nx;                         // evaluates the face between elements ic-1 and ic
  const bool virt = ic == g.nx;                              // ... as the (wrapped / clamped) element behind the last column
  const int start = (ic0 - 1) & ~1;                          // first box column: even => 16-byte aligned
  const int col = ic - start;                                // own column inside the box
  const int icL = bc_index(P.bc, -1, g.nyg), icR = bc_index(P.bc, g.nx, g.nyg);   // the reference wraps x with ny (:1338)
  const uint32_t slot0 = tma::smem_u32(dgm_smem), bars = slot0 + 2 * REGION_B;
  const double* S[2] = {reinterpret_cast<const double*>(dgm_smem), reinterpret_cast<const double*>(dgm_smem + REGION_B)};
  uint32_t ph[2] = {0, 0};                                   // phase parity of the two slot barriers
  auto wait_slot = [&](int k) { tma::mbar_wait(bars + 8 * k, ph[k]); ph[k] ^= 1; };
  auto arm_slot = [&](int k, int row) {                      // lane 0: slot k <- row `row`
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma::mbar_expect_tx(bars + 8 * k, REGION_B);
    tma::load_3d(slot0 + k * REGION_B, &m_in, start, row, 0, bars + 8 * k);
  };
  if (lane == 0) {
    tma::mbar_init(bars, 1); tma::mbar_init(bars + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    arm_slot(0, y_nb(g, P.bc, j0 - 1));                      // row below the strip: bottom face of its first row
    arm_slot(1, j0);
  }
  __syncwarp();
  const bool a0_in = C.A0 == in, a1_in = C.A1 == in;
  int s = 1;                                                 // slot of the own row
  double FB[M][4];                                           // bottom face flux, carried up the strip
  for (int jc = j0; jc < j1; ++jc) {
    const double* own = S[s];
    const double* other = S[s ^ 1];                          // row below (first row of the strip), then the row above
    const size_t erow = (size_t)jc * g.nx;
    double d[4][M][M], acc[4][M][M];
    double FL[M][4], FR[M][4];
    if (jc == j0) wait_slot(s);                              // later own rows were waited for as "row above"
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (virt) load_var<M>(in, g, v, erow + icR, d[v]); else march_from_smem<M>(own, v, col, d[v]);
    }
    if (jc == j0) {                                          // first row of the strip: its bottom face
      wait_slot(s ^ 1);
      double tn[M][4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        double t1[M], dn[M][M];
        trace1<M, 2>(d[v], B, t1);
#pragma unroll
        for (int q = 0; q < M; ++q) FB[q][v] = t1[q];
        march_from_smem<M>(other, v, col, dn);
        trace1<M, 3>(dn, B, t1);
#pragma unroll
        for (int q = 0; q < M; ++q) tn[q][v] = t1[q];
      }
      if (owner) face_flux_from_traces<M, 2, ANYFLUX>(P, FB, tn);
      __syncwarp();                                          // every lane has read the row below
      if (lane == 0) arm_slot(s ^ 1, y_nb(g, P.bc, jc + 1)); // ... its slot now receives the row above
    }
    // ---- left face
    {
      double tn[M][4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        double t1[M], dn[M][M];
        trace1<M, 0>(d[v], B, t1);
#pragma unroll
        for (int q = 0; q < M; ++q) FL[q][v] = t1[q];
        if (ic == 0) load_var<M>(in, g, v, erow + icL, dn); else march_from_smem<M>(own, v, col - 1, dn);
        trace1<M, 1>(dn, B, t1);
#pragma unroll
        for (int q = 0; q < M; ++q) tn[q][v] = t1[q];
      }
      if (face_lane) face_flux_from_traces<M, 0, ANYFLUX>(P, FL, tn);
    }
    // ---- right face = left face of the next lane
#pragma unroll
    for (int q = 0; q < M; ++q)
#pragma unroll
      for (int v = 0; v < 4; ++v) FR[q][v] = __shfl_down_sync(0xffffffffu, FL[q][v], 1);
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int a = 0; a < M; ++a)
#pragma unroll
        for (int b = 0; b < M; ++b) acc[v][a][b] = 0.0;
    face_accum<M, 0>(B, FL, acc);
    face_accum<M, 1>(B, FR, acc);
    face_accum<M, 2>(B, FB, acc);
    // ---- top face (row above in the other slot); its flux is the next row's bottom flux
    wait_slot(s ^ 1);
    if (owner) {
      double tn[M][4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        double t1[M], dn[M][M];
        trace1<M, 3>(d[v], B, t1);
#pragma unroll
        for (int q = 0; q < M; ++q) FB[q][v] = t1[q];
        march_from_smem<M>(other, v, col, dn);
        trace1<M, 2>(dn, B, t1);
#pragma unroll
        for (int q = 0; q < M; ++q) tn[q][v] = t1[q];
      }
      face_flux_from_traces<M, 3, ANYFLUX>(P, FB, tn);
      face_accum<M, 3>(B, FB, acc);
      MarchRk<M> rk{in, g, erow + ic, own + col, a0_in, a1_in};
      dg_stage_rest<M>(rk, RegModes<M>{d}, acc, C, out, gx, gy, fz, g, P, B, ctrl, apply_onp, erow + ic);
    }
    __syncwarp();                                            // the own row is no longer needed
    if (jc + 1 < j1 && lane == 0) arm_slot(s, y_nb(g, P.bc, jc + 2));   // row above of the next iteration
    s ^= 1;
  }
}

}}  // namespace wb::dg
