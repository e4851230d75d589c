// k_dg_stage_split: the fused DG 2D RK-stage kernel with the ELEMENT SPLIT OVER FOUR THREADS (one per conserved variable).
//
// Why: round 1's kernel (k_dg_stage_tma, since removed) kept one element per thread -- 36 modes + 36 accumulators + the nodal fluxes in one thread = 255
// registers, ~0.5 KB of spills, 8 warps per SM, 41 % of the HBM peak (profiles/r1_dg2d_tma_kernel_final.txt); it also evaluated
// every face from both sides.  Everything in the stage that is LINEAR (traces, nodal values, volume / edge / source integrals,
// RK combination) is independent per variable, so here
//   * a block is 4 warps = 32 consecutive elements of a row x 4 variables: warp v owns variable v, lane l owns column ic0+l.
//     A thread holds 9 accumulators + 9 nodal values of ONE variable (M = 3): <= 128 registers, no spills, 16 warps per SM;
//   * the POINTWISE (non-linear) parts -- primitive variables at the M*M volume nodes, the numerical flux at the M points of a
//     face -- are "work items" dealt to all 128 threads through shared memory: 9 volume nodes + 3 left-face points + 3 top-face
//     points per element, + the one x face behind the last column of the block;
//   * every face flux is evaluated ONCE: the block marches up a strip of rows; the top-face flux of row j is the bottom-face
//     flux of row j+1 (carried in registers by the thread that needs it), the right-face flux of column c is the left-face
//     flux of column c+1;
//   * rows travel HBM -> shared memory once per strip as TMA tensor boxes (36 columns x 1 row x all planes, two slots,
//     one mbarrier each); the slot of row j is re-armed with row j+2 as soon as the traces and nodal values of row j exist.
// The arithmetic (operation order of every sum, the LLF routine with the low side first, the RK combination, the 'ONP' test)
// is that of k_dg_stage_fast, so the results are BIT-IDENTICAL to it (tests/test_dg2d_gpu.py).
//
// Per iteration (row j of the strip):
//   A  variable threads: own row -> left/right/top traces + nodal values (shared memory); row j+1 -> bottom trace
//      ---- barrier; thread 0 re-arms the slot of row j with row j+2
//   B  item threads: prim() and the flux pieces that need the conserved values at the nodes (in place of the nodal values),
//      LLF at the left-face points (in place of the left traces) and at the top-face points (in place of the top traces)
//      ---- barrier
//   C  variable threads: edge integrals (left, right, bottom, top), fluxes of the own variable from the primitives, volume
//      and source integrals, RK combination, sufficient test of 'ONP' (needs all four variables: one more exchange), stores
// Reference: 2d/benchmark_2d_dg.f90:1137-1479 (compute_update), :683-707 (RK combination), 2d/limiters.f90:478-654 ('ONP').
// Slab boundary-row launches also store their rows into the neighbours' ghost rows (peer memory, StageCoef::peer_*).
// Included by dg2d_split.cu after dg2d_fast.cuh and dg2d_tma.cuh (DGT_W, tma:: wrappers).  Needs nx % 32 == 0.
#pragma once

namespace wb { namespace dg {

template <int M>
struct SplitLayout {                                   // shared memory, in doubles
  static constexpr int NM = M * M, NP = 4 * NM;
  static constexpr int SLOT_D = NP * DGT_W;            // one TMA row slot
  static constexpr int TW = 34;                        // columns of the x-trace arrays (33 used: columns -1..31 / 0..32)
  static constexpr int NS = 7;                         // slots per volume node: nodal values (0..3) -> both fluxes (0..6: the two
                                                       // fluxes share rho*vx*vy, slot 2)
  static constexpr int OFF_UB = 2 * SLOT_D;            // [NM nodes][NS][32]
  static constexpr int OFF_TL = OFF_UB + NM * NS * 32; // [M][4][TW] left traces of columns 0..32 -> left-face fluxes
  static constexpr int OFF_TR = OFF_TL + M * 4 * TW;   // [M][4][TW] right traces of columns -1..31 (index c+1)
  static constexpr int OFF_TT = OFF_TR + M * 4 * TW;   // [M][4][32] top traces of the row -> top-face fluxes (the bottom-face
                                                       // fluxes are last row's top-face fluxes: carried in registers)
  static constexpr int OFF_TB = OFF_TT + M * 4 * 32;   // [M][4][32] bottom traces of the row above (phases A, B) ...
  static constexpr int OFF_LIM = OFF_TB;               // ... and [8][32] mean and mode bound of every variable ('ONP' test, end of
                                                       // phase C: TB is dead by then and the vote barrier closes the reads)
  static constexpr int OFF_SW = OFF_TB + (M * 4 * 32 > 8 * 32 ? M * 4 * 32 : 8 * 32);
                                                       // [NM nodes][3][32] gravity (g1, g2) -> sources (kernels with a source term)
  static constexpr int OFF_BAR_SRC = OFF_SW + NM * 3 * 32, OFF_BAR_NOSRC = OFF_SW;      // 2 mbarriers + the step's dt
  template <bool SRC> static constexpr int bytes() { return ((SRC ? OFF_BAR_SRC : OFF_BAR_NOSRC) + 3) * 8; }
  // work items of phase B: [0, NM) volume nodes, [NM, NM+M) left-face points, [NM+M, NM+2M) top-face points,
  // NM+2M: the x face behind the last column (M points, lanes 0..M-1)
  static constexpr int NTYPES = NM + 2 * M + 1;
};

// Item schedule of phase B.  Items: M+1 x-face items (q = 0..M-1: point q of the left face of the warp's 32 columns; q = M:
// the M points of the face behind the last column, lanes 0..M-1), M y-face items, M*M volume nodes.  Warp w works through
// the contiguous ranges [x0,x1), [y0,y1), [n0,n1) packed into one int (4 bits each, from bit 0).  Measured cost in issue
// slots (an FP64 instruction takes two): volume node ~65, face item ~175 (two prim() + two sound speeds + LLF).
// M = 3: w0: x 0,1 + nodes 0,1 | w1: x 2, y 0 + nodes 2,3 | w2: y 1 + nodes 4..7 | w3: x 3 (extra), y 2 + node 8
template <int M>
__device__ __forceinline__ unsigned split_schedule(int w, int variant) {
  constexpr int NM = M * M;
  auto pack = [](int x0, int x1, int y0, int y1, int n0, int n1) {
    return (unsigned)(x0 | (x1 << 4) | (y0 << 8) | (y1 << 12) | (n0 << 16) | (n1 << 21));
  };
  if (M == 3) {
    if (variant == 1) return w == 0 ? pack(0, 2, 0, 0, 0, 1) : w == 1 ? pack(2, 3, 0, 1, 1, 2) : w == 2 ? pack(3, 3, 1, 2, 2, 7) : pack(3, 4, 2, 3, 7, 9);
    if (variant == 2) return w == 0 ? pack(0, 2, 0, 0, 0, 0) : w == 1 ? pack(2, 3, 0, 1, 0, 0) : w == 2 ? pack(3, 3, 1, 2, 0, 6) : pack(3, 4, 2, 3, 6, 9);
    if (variant == 3) return w == 0 ? pack(0, 2, 0, 0, 0, 3) : w == 1 ? pack(2, 3, 0, 1, 3, 5) : w == 2 ? pack(3, 3, 1, 2, 5, 8) : pack(3, 4, 2, 3, 8, 9);
    return w == 0 ? pack(0, 2, 0, 0, 0, 2) : w == 1 ? pack(2, 3, 0, 1, 2, 4) : w == 2 ? pack(3, 3, 1, 2, 4, 8) : pack(3, 4, 2, 3, 8, 9);
  }
  return pack(w * (M + 1) / 4, (w + 1) * (M + 1) / 4, w * M / 4, (w + 1) * M / 4, w * NM / 4, (w + 1) * NM / 4);
}

template <int M, int FACE>
__device__ __forceinline__ void face_accum1(const FastBasis& B, const double (&F)[M], double (&acc)[M][M]) {
  double s[M];
#pragma unroll
  for (int n = 0; n < M; ++n) {
    double a1 = 0.0;
#pragma unroll
    for (int q = 0; q < M; ++q)
      if (!zP<M>(q, n)) a1 = fma(F[q], B.Pwh[q][n], a1);               // F = 2 * flux, Pwh = Pw / 2
    s[n] = a1;
  }
#pragma unroll
  for (int a = 0; a < M; ++a)
#pragma unroll
    for (int b = 0; b < M; ++b) {
      if (FACE == 0) acc[a][b] = fma(B.Em[a], s[b], acc[a][b]);     // + e2
      if (FACE == 1) acc[a][b] = fma(-B.Ep[a], s[b], acc[a][b]);    // - e1
      if (FACE == 2) acc[a][b] = fma(B.Em[b], s[a], acc[a][b]);     // + e4
      if (FACE == 3) acc[a][b] = fma(-B.Ep[b], s[a], acc[a][b]);    // - e3
    }
}

#ifndef DGS_MINB
#define DGS_MINB 4
#endif
#ifndef DGS_MINB_OUT2
#define DGS_MINB_OUT2 4
#endif
#ifndef DGS_MINB_SRC2
#define DGS_MINB_SRC2 3      // the two-result launch with a source term: 168 registers without spills beat 128 with 100 B of them
#endif

template <int M, bool ANYFLUX, bool SRC, bool OUT2>
__global__ void __launch_bounds__(128, (M <= 3 ? ((SRC && OUT2) ? DGS_MINB_SRC2 : OUT2 ? DGS_MINB_OUT2 : DGS_MINB) : 2))
k_dg_stage_split(const __grid_constant__ CUtensorMap m_in, const double* __restrict__ in, StageCoef C, double* __restrict__ out,
                 const double* __restrict__ gx, const double* __restrict__ gy, const unsigned char* __restrict__ fz, DgGrid g,
                 DgPhys P, const __grid_constant__ FastBasis B, const DgCtrl* __restrict__ ctrl, int apply_onp, int rows_sched,
                 int row_begin, int row_end) {
  using L = SplitLayout<M>;
  constexpr int NM = L::NM, NS = L::NS;
  extern __shared__ __align__(128) unsigned char dgs_smem[];
  double* sm = reinterpret_cast<double*>(dgs_smem);
  const int lane = threadIdx.x & 31, v = threadIdx.x >> 5;
  const int ic0 = blockIdx.x * 32;
  const int rows = rows_sched & 0xffff;                      // strip height; bits 16.. = phase-B schedule variant
  const int j0 = row_begin + blockIdx.y * rows, j1 = min(j0 + rows, row_end);
  if (j0 >= j1) return;
  if (ctrl->skip) {
    for (int jc = j0; jc < j1; ++jc) {
      const size_t e = (size_t)jc * g.nx + ic0 + lane;
      double* pp = jc == 1 ? C.peer_lo : jc == g.ny - 2 ? C.peer_hi : nullptr;       // the handed-through rows travel as well
      const size_t pne = jc == 1 ? C.peer_lo_ne : C.peer_hi_ne;
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        const double x = PL(in, g, v, m)[e];
        PL(out, g, v, m)[e] = x;
        if (pp) pp[((size_t)v * NM + m) * pne + ic0 + lane] = x;
      }
    }
    return;
  }
  double* UB = sm + L::OFF_UB;
  double* TL = sm + L::OFF_TL;
  double* TR = sm + L::OFF_TR;
  double* TT = sm + L::OFF_TT;
  double* TB = sm + L::OFF_TB;
  double* LIM = sm + L::OFF_LIM;
  double* SW = sm + L::OFF_SW;
  const uint32_t slot0 = tma::smem_u32(sm), bars = tma::smem_u32(sm + (SRC ? L::OFF_BAR_SRC : L::OFF_BAR_NOSRC));
  constexpr uint32_t SLOT_BYTES = L::SLOT_D * 8;
  auto arm_slot = [&](int k, int row) {                      // one thread: slot k <- row `row`, columns ic0-2 .. ic0+33
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tma::mbar_expect_tx(bars + 8 * k, SLOT_BYTES);
    tma::load_3d(slot0 + k * SLOT_BYTES, &m_in, ic0 - 2, row, 0, bars + 8 * k);      // even column: 16-byte aligned
  };
  if (threadIdx.x == 0) {
    tma::mbar_init(bars, 1); tma::mbar_init(bars + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    arm_slot(0, y_nb(g, P.bc, j0 - 1));                      // row below the strip: only its top trace is needed
    arm_slot(1, j0);
    sm[(SRC ? L::OFF_BAR_SRC : L::OFF_BAR_NOSRC) + 2] = ctrl->dt;      // read from shared memory once per row: no global-load latency there
  }
  __syncthreads();
  // Iteration `it` handles row j = j0-1+it (it = 0 is the prologue: top face of the row below the strip only).  Row j lies
  // in slot it&1 and row j+1 in the other one; the k-th load of a slot completes phase k&1 of its mbarrier, and the wait
  // for row j+1 in iteration `it` is load number (it+1)>>1 of slot (it+1)&1.
  tma::mbar_wait(bars, 0);
  const int nit = j1 - j0 + 1;
  const unsigned sched = split_schedule<M>(v, rows_sched >> 16);
  double Fbot[M];                                            // own variable's fluxes through the bottom face of the current row
#pragma unroll
  for (int q = 0; q < M; ++q) Fbot[q] = 0.0;
#pragma unroll 1
  for (int it = 0; it < nit; ++it) {
    const int j = j0 - 1 + it;
    const bool full = it > 0;
    const int par = it & 1;
    double* TTp = TT;
    // RK operands of the NEXT row: ask L2 for them now (a plane's 32 columns are two 128-byte lines: lane l of warp v touches
    // line l&1 of plane v*NM + l/2), so that phase C of the next iteration finds them in L2 instead of waiting for DRAM
    // (ncu: long_scoreboard on the first use of A0 was 5-15 % of the stall samples).  Per-lane prefetch.global.L2, not the
    // bulk form: that one takes a uniform address and costs a loop over the lanes.
    if (lane < 2 * NM && j + 1 < j1) {
      const size_t po = ((size_t)v * NM + (lane >> 1)) * g.ne + (size_t)(j + 1) * g.nx + ic0 + (lane & 1) * 16;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(C.A0 + po));
      if (C.na >= 2 && C.A1 != in) asm volatile("prefetch.global.L2 [%0];" ::"l"(C.A1 + po));
      if (OUT2) asm volatile("prefetch.global.L2 [%0];" ::"l"(C.B1 + po));
      if (SRC && v == 0 && P.source == 2 && !P.gsep) {       // the gravity field of the next row (read by the node items of phase B)
        const size_t pg = (size_t)(lane >> 1) * g.ne + (size_t)(j + 1) * g.nx + ic0 + (lane & 1) * 16;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(gx + pg));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(gy + pg));
      }
    }
    // gravity field of this row: warp v < M fetches the M nodes qx = v, qy = 0..M-1 of its column now (planes qy*M + qx of both
    // arrays: one base address + a constant plane stride) and parks them in the source slots at the end of phase A, where the
    // node items of phase B find them -- the L2 latency hides behind phase A.  Separable field (DgPhys::gsep): the same numbers
    // from the one row of gx (plane qx, row `grow`: ONE value for all qy) and the one column of gy (plane qy*M, column 0).
    constexpr int NG = M;
    double ga[SRC ? NG : 1], gb[SRC ? NG : 1];
    if (SRC && full && P.source == 2 && v < M) {
      const size_t strideq = (size_t)M * g.ne;
      if (P.gsep) {
        const double gx1 = gx[(size_t)v * g.ne + (size_t)P.grow * g.nx + ic0 + lane];
        const double* py = gy + (size_t)j * g.nx;
#pragma unroll
        for (int i = 0; i < NG; ++i) { ga[SRC ? i : 0] = gx1; gb[SRC ? i : 0] = py[i * strideq]; }
      } else {
        const size_t o = (size_t)v * g.ne + (size_t)j * g.nx + ic0 + lane;
#pragma unroll
        for (int i = 0; i < NG; ++i) { ga[SRC ? i : 0] = gx[o + i * strideq]; gb[SRC ? i : 0] = gy[o + i * strideq]; }
      }
    }
    // ------------------------------------------------------------------ phase A
    {
      const double* Sc = sm + par * L::SLOT_D + lane + 2;    // own column of row j
      double d[M][M], t1[M];
#pragma unroll
      for (int jm = 0; jm < M; ++jm)
#pragma unroll
        for (int im = 0; im < M; ++im) d[im][jm] = Sc[(v * NM + jm * M + im) * DGT_W];
      trace1<M, 3>(d, B, t1);
#pragma unroll
      for (int q = 0; q < M; ++q) TTp[(q * 4 + v) * 32 + lane] = t1[q];
      if (full) {
        trace1<M, 0>(d, B, t1);
#pragma unroll
        for (int q = 0; q < M; ++q) TL[(q * 4 + v) * L::TW + lane] = t1[q];
        trace1<M, 1>(d, B, t1);
#pragma unroll
        for (int q = 0; q < M; ++q) TR[(q * 4 + v) * L::TW + lane + 1] = t1[q];
#pragma unroll
        for (int qx = 0; qx < M; ++qx) {
          double a[M];                                         // P[q][0] = 1: the sums start with their first term
#pragma unroll
          for (int jm = 0; jm < M; ++jm) {
            double s = d[0][jm];
#pragma unroll
            for (int im = 1; im < M; ++im)
              if (!zP<M>(qx, im)) s = fma(d[im][jm], B.P[qx][im], s);
            a[jm] = s;
          }
#pragma unroll
          for (int qy = 0; qy < M; ++qy) {
            double s = a[0];
#pragma unroll
            for (int jm = 1; jm < M; ++jm)
              if (!zP<M>(qy, jm)) s = fma(a[jm], B.P[qy][jm], s);
            UB[((qx * M + qy) * NS + v) * 32 + lane] = s;
          }
        }
        // the two columns beside the block: right trace of column -1 (lane 0), left trace of column 32 (lane 31)
        if (lane == 0 || lane == 31) {
          // at the edge of the domain the x neighbour is not in the box; the reference wraps x with ny (:1338)
          const bool glob = lane == 0 ? ic0 == 0 : ic0 + 32 == g.nx;
          if (glob) {
            const size_t en = (size_t)j * g.nx + (lane == 0 ? bc_index(P.bc, -1, g.nyg) : bc_index(P.bc, g.nx, g.nyg));
#pragma unroll
            for (int jm = 0; jm < M; ++jm)
#pragma unroll
              for (int im = 0; im < M; ++im) d[im][jm] = PL(in, g, v, jm * M + im)[en];
          } else {
            const int off = lane == 0 ? -1 : 1;
#pragma unroll
            for (int jm = 0; jm < M; ++jm)
#pragma unroll
              for (int im = 0; im < M; ++im) d[im][jm] = Sc[(v * NM + jm * M + im) * DGT_W + off];
          }
          if (lane == 0) {
            trace1<M, 1>(d, B, t1);
#pragma unroll
            for (int q = 0; q < M; ++q) TR[(q * 4 + v) * L::TW + 0] = t1[q];
          } else {
            trace1<M, 0>(d, B, t1);
#pragma unroll
            for (int q = 0; q < M; ++q) TL[(q * 4 + v) * L::TW + 32] = t1[q];
          }
        }
      }
      tma::mbar_wait(bars + 8 * (par ^ 1), ((it + 1) >> 1) & 1);      // row j+1
      const double* Sn = sm + (par ^ 1) * L::SLOT_D + lane + 2;
#pragma unroll
      for (int jm = 0; jm < M; ++jm)
#pragma unroll
        for (int im = 0; im < M; ++im) d[im][jm] = Sn[(v * NM + jm * M + im) * DGT_W];
      trace1<M, 2>(d, B, t1);
#pragma unroll
      for (int q = 0; q < M; ++q) TB[(q * 4 + v) * 32 + lane] = t1[q];
    }
    if (SRC && full && P.source == 2 && v < M) {
#pragma unroll
      for (int i = 0; i < NG; ++i) {                         // node k = qx*M + qy = v*M + i
        SW[((v * M + i) * 3) * 32 + lane] = ga[SRC ? i : 0];
        SW[((v * M + i) * 3 + 1) * 32 + lane] = gb[SRC ? i : 0];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && j + 1 < j1) arm_slot(par, y_nb(g, P.bc, j + 2));   // row j is consumed
    // ------------------------------------------------------------------ phase B
    if (full) {
#pragma unroll 1
      for (int q = sched & 15; q < ((sched >> 4) & 15); ++q) {      // x face between columns c-1 and c, point qq
        const int qq = q == M ? lane : q, c = q == M ? 32 : lane;
        if (q == M && lane >= M) break;
        const double* pl = TR + (qq * 4) * L::TW + c;          // right trace of column c-1
        double* pr = TL + (qq * 4) * L::TW + c;                // left trace of column c
        const double ul[4] = {pl[0], pl[L::TW], pl[2 * L::TW], pl[3 * L::TW]};
        const double ur[4] = {pr[0], pr[L::TW], pr[2 * L::TW], pr[3 * L::TW]};
        double F[4];
        fastm::llf<1, ANYFLUX, true>(P, ul, ur, F);
        pr[0] = F[0]; pr[L::TW] = F[1]; pr[2 * L::TW] = F[2]; pr[3 * L::TW] = F[3];
      }
    }
#pragma unroll 1
    for (int q = (sched >> 8) & 15; q < ((sched >> 12) & 15); ++q) {  // y face between rows j and j+1, point q
      double* pl = TTp + (q * 4) * 32 + lane;                  // top trace of row j
      const double* pr = TB + (q * 4) * 32 + lane;             // bottom trace of row j+1
      const double ul[4] = {pl[0], pl[32], pl[64], pl[96]};
      const double ur[4] = {pr[0], pr[32], pr[64], pr[96]};
      double F[4];
      fastm::llf<2, ANYFLUX, true>(P, ul, ur, F);
      pl[0] = F[0]; pl[32] = F[1]; pl[64] = F[2]; pl[96] = F[3];
    }
    if (full) {
      // volume node: slots (rho, mx, my, E) -> both fluxes of compute_flux (:919-944), so that phase C only picks the
      // two numbers of its variable; kernels with a source term also keep (w0, vx, vy)
#pragma unroll 1
      for (int k = (sched >> 16) & 31; k < ((sched >> 21) & 31); ++k) {
        double* p = UB + (k * NS) * 32 + lane;
        // kernels with a source term: the source of every variable at this node, evaluated ONCE here (get_source :1558-1576:
        // (0, w0 g1, w0 g2, w0 (vx g1 + vy g2)); get_adv_source :1579-1596: (-rho, 0, 0, 0)) -- phase C picks its variable's.
        // (g1, g2) were fetched at the start of the row and wait in the slots that take the sources.
        double g1 = 0.0, g2 = 0.0;
        if (SRC && P.source == 2) { g1 = SW[(k * 3) * 32 + lane]; g2 = SW[(k * 3 + 1) * 32 + lane]; }
        const double u0 = p[0], u1 = p[32], u2 = p[64], u3 = p[96];
        const fastm::Prim w = fastm::prim(P, u0, u1, u2, u3);
        const double t = w.w0 * w.vx * w.vy, Ep = u3 + w.p;
        p[0] = w.w0 * w.vx; p[32] = fma(w.vx, u1, w.p); p[64] = t; p[96] = w.vx * Ep;      // x flux: slots 0..3
        p[128] = w.w0 * w.vy; p[160] = fma(w.vy, u2, w.p); p[192] = w.vy * Ep;               // y flux: slots 4, 2, 5, 6
        if (SRC) {
          double* q = SW + (k * 3) * 32 + lane;
          if (P.source == 2) { q[0] = w.w0 * g1; q[32] = w.w0 * g2; q[64] = w.w0 * fma(w.vx, g1, w.vy * g2); }
          else q[0] = u0;
        }
      }
    }
    __syncthreads();
    if (!full) {                                             // row below the strip: only its top-face fluxes are needed
#pragma unroll
      for (int q = 0; q < M; ++q) Fbot[q] = TT[(q * 4 + v) * 32 + lane];
      continue;
    }
    // ------------------------------------------------------------------ phase C
    const size_t e = (size_t)j * g.nx + ic0 + lane;
    const size_t pe = (size_t)v * NM * g.ne + e;             // element of plane (v, 0)
    // RK operands: asked for now, used after the volume integrals (their DRAM latency hides behind phase C's arithmetic)
    double rk0[M][M], rk1[M][M], rk2[OUT2 ? M : 1][OUT2 ? M : 1];
#pragma unroll
    for (int b = 0; b < M; ++b)
#pragma unroll
      for (int a = 0; a < M; ++a) {
        rk0[a][b] = C.A0[pe + (size_t)(b * M + a) * g.ne];
        if (C.na >= 2) rk1[a][b] = C.A1[pe + (size_t)(b * M + a) * g.ne];
        if (OUT2) rk2[OUT2 ? a : 0][OUT2 ? b : 0] = C.B1[pe + (size_t)(b * M + a) * g.ne];
      }
    double acc[M][M];
#pragma unroll
    for (int a = 0; a < M; ++a)
#pragma unroll
      for (int b = 0; b < M; ++b) acc[a][b] = 0.0;
    {
      double F[M];
#pragma unroll
      for (int q = 0; q < M; ++q) F[q] = TL[(q * 4 + v) * L::TW + lane];
      face_accum1<M, 0>(B, F, acc);
#pragma unroll
      for (int q = 0; q < M; ++q) F[q] = TL[(q * 4 + v) * L::TW + lane + 1];
      face_accum1<M, 1>(B, F, acc);
      face_accum1<M, 2>(B, Fbot, acc);                       // bottom face = the top face of the row below
#pragma unroll
      for (int q = 0; q < M; ++q) F[q] = TT[(q * 4 + v) * 32 + lane];
      face_accum1<M, 3>(B, F, acc);
#pragma unroll
      for (int q = 0; q < M; ++q) Fbot[q] = F[q];
    }
    {
      // Per quadrature row qy: fluxes of the own variable at its M nodes, their contraction over qx, and the update of the
      // M*M accumulators -- the same sums in the same order as k_dg_stage_fast (vol: s = acc, then qy ascending; source:
      // s = 0, then qy ascending), with M instead of M*M fluxes alive.
      const int f2_slot = v == 0 ? 4 : v == 1 ? 2 : v + 3;   // the y flux of variable v (see the node items)
#pragma unroll
      for (int qy = 0; qy < M; ++qy) {
        double f1[M], f2[M];
#pragma unroll
        for (int qx = 0; qx < M; ++qx) {
          const double* p = UB + ((qx * M + qy) * NS) * 32 + lane;
          f1[qx] = p[v * 32]; f2[qx] = p[f2_slot * 32];
        }
#pragma unroll
        for (int a = 0; a < M; ++a) {
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int qx = 0; qx < M; ++qx) {
            if (a > 0 && !zD<M>(qx, a)) s1 = fma(f1[qx], B.dPw[qx][a], s1);   // exact zeros of the tables: terms dropped
            if (!zP<M>(qx, a)) s2 = fma(f2[qx], B.Pw[qx][a], s2);
          }
#pragma unroll
          for (int b = 0; b < M; ++b) {
            if (b > 0 && !zD<M>(qy, b)) acc[a][b] = fma(s2, B.dPw[qy][b], acc[a][b]);
            if (a > 0 && !zP<M>(qy, b)) acc[a][b] = fma(s1, B.Pw[qy][b], acc[a][b]);
          }
        }
      }
      if (SRC) {
        // Source integral (:1403-1447), its own sums (s = 0, then qy ascending) added to the volume integral at the end.
        // gravity (source 2): variable v > 0 finds its source in slot v-1; advection sink (source 3): variable 0 takes -rho
        // from slot 0.  A warp is one variable, so the test is warp-uniform: the variables without a source skip the sums
        // and add the +0.0 the sums would have produced (-0.0 + 0.0 = +0.0: same bits as fma(scale, +0.0, acc)).
        const bool src_on = (P.source == 2) ? v != 0 : v == 0;
        if (src_on) {
          // the sink's minus sign rides on the final scale: negation commutes exactly with every product and sum below
          const double src_scale = (P.source == 2) ? 0.5 * P.dx : -(0.5 * P.dx);
          const double* sw = SW + ((P.source == 2) ? v - 1 : 0) * 32 + lane;
          double sv[M][M];
#pragma unroll
          for (int a = 0; a < M; ++a)
#pragma unroll
            for (int b = 0; b < M; ++b) sv[a][b] = 0.0;
#pragma unroll
          for (int qy = 0; qy < M; ++qy) {
            double S[M];
#pragma unroll
            for (int qx = 0; qx < M; ++qx) S[qx] = sw[((qx * M + qy) * 3) * 32];
#pragma unroll
            for (int a = 0; a < M; ++a) {
              double s3 = 0.0;
#pragma unroll
              for (int qx = 0; qx < M; ++qx)
                if (!zP<M>(qx, a)) s3 = fma(S[qx], B.Pw[qx][a], s3);
#pragma unroll
              for (int b = 0; b < M; ++b)
                if (!zP<M>(qy, b)) sv[a][b] = fma(s3, B.Pw[qy][b], sv[a][b]);
            }
          }
#pragma unroll
          for (int a = 0; a < M; ++a)
#pragma unroll
            for (int b = 0; b < M; ++b) acc[a][b] = fma(src_scale, sv[a][b], acc[a][b]);
        } else {
#pragma unroll
          for (int a = 0; a < M; ++a)
#pragma unroll
            for (int b = 0; b < M; ++b) acc[a][b] = acc[a][b] + 0.0;
        }
      }
    }
    // ---- dudt scaling (:1449-1466), RK combination (:683-707).  c0*a0 with c0 == 1 is a0 exactly, so the first stage needs
    //      no special case; the frozen modes of special_boundary_conditions (:1481-1514) are a separate, rarely taken pass.
    const double dt = sm[(SRC ? L::OFF_BAR_SRC : L::OFF_BAR_NOSRC) + 2];
#pragma unroll
    for (int b = 0; b < M; ++b)
#pragma unroll
      for (int a = 0; a < M; ++a) acc[a][b] = (0.5 * P.oneoverdx) * acc[a][b];
    if (fz) {
#pragma unroll
      for (int b = 0; b < M; ++b)
#pragma unroll
        for (int a = 0; a < M; ++a)
          if (fz[(size_t)(b * M + a) * g.ne + e]) acc[a][b] = 0.0;
    }
    double W[OUT2 ? M : 1][OUT2 ? M : 1];                    // second result without its k3*out term
    if (OUT2) {
      // SSPRK(5,4) stage 4 emits w5 as well (:700-704): B0 is A0 (delta_u) and the stage input is A1 there -- the values
      // already asked for are used again instead of being loaded twice
      const double kedt = C.ke * dt;
      const bool b0_is_a0 = C.B0 == C.A0, in_is_a1 = C.na >= 2 && in == C.A1;
#pragma unroll
      for (int b = 0; b < M; ++b)
#pragma unroll
        for (int a = 0; a < M; ++a) {
          const size_t pm = pe + (size_t)(b * M + a) * g.ne;
          double r2 = C.k0 * (b0_is_a0 ? rk0[a][b] : C.B0[pm]);
          r2 = fma(C.k1, rk2[OUT2 ? a : 0][OUT2 ? b : 0], r2);
          r2 = fma(C.k2, in_is_a1 ? rk1[a][b] : in[pm], r2);
          W[OUT2 ? a : 0][OUT2 ? b : 0] = fma(kedt, acc[a][b], r2);
        }
    }
    {
      const double cdt = C.cd * dt;
      if (C.na >= 2) {
#pragma unroll
        for (int b = 0; b < M; ++b)
#pragma unroll
          for (int a = 0; a < M; ++a) acc[a][b] = fma(cdt, acc[a][b], fma(C.c1, rk1[a][b], C.c0 * rk0[a][b]));
      } else {
#pragma unroll
        for (int b = 0; b < M; ++b)
#pragma unroll
          for (int a = 0; a < M; ++a) acc[a][b] = fma(cdt, acc[a][b], C.c0 * rk0[a][b]);
      }
    }
    // ---- 'ONP' (2d/limiters.f90:478-654): sufficient test on bounds of the point values (see positivity_fast)
    if (apply_onp && M > 1) {
      double R = 0.0;
#pragma unroll
      for (int a = 0; a < M; ++a)
#pragma unroll
        for (int b = 0; b < M; ++b)
          if (a != 0 || b != 0) R = fma(fabs(acc[a][b]), B.EpEp[a][b], R);
      LIM[v * 32 + lane] = acc[0][0];
      LIM[(4 + v) * 32 + lane] = R;
      __syncthreads();
      const double ua0 = LIM[lane], ua1 = LIM[32 + lane], ua2 = LIM[64 + lane], ua3 = LIM[96 + lane];
      const double R0 = LIM[128 + lane], R1 = LIM[160 + lane], R2 = LIM[192 + lane], R3 = LIM[224 + lane];
      const double rho_lo = ua0 - R0;
      const double mx_hi = fabs(ua1) + R1, my_hi = fabs(ua2) + R2, E_lo = ua3 - R3;
      const double margin = 1e-9 * (fabs(ua0) + R0 + fabs(ua3) + R3);
      bool ok = false;
      if (rho_lo > P.eps + margin && rho_lo > (double)10e-10f) {
        const double p_lo = P.gm1a * (E_lo - 0.5 * (mx_hi * mx_hi + my_hi * my_hi) / rho_lo);
        ok = p_lo > P.eps + margin;
      }
      if (__syncthreads_or(!ok)) {
        // Rare: some element of the block fails the test and needs the point evaluations of compute_positivity.  They
        // are done by the element's four threads together -- the density limiter by thread 0, the point values of the
        // pressure limiter by every thread for its variable, exchanged through the (dead) node buffer, one point family at
        // a time; same evaluations and the same minima as positivity_slow (dg2d_fast.cuh), so the same bits.
        constexpr int GLL = (2 * (M - 1) + 3) / 2;           // compute_set :438-475
        double* PT = UB;                                     // [M*GLL points][4][32]
        if (v == 0 && !ok) {
          double p_min = 1e300;
#pragma unroll 1
          for (int q = 0; q < M; ++q)
#pragma unroll 1
            for (int r = 0; r < GLL; ++r) {
              p_min = fmin(p_min, eval_at<M>(acc, B.Pg[r], B.P[q]));     // GLL in x, GL in y
              p_min = fmin(p_min, eval_at<M>(acc, B.P[q], B.Pg[r]));     // GL in x, GLL in y
            }
          const double theta = fmin(fabs((ua0 - P.eps) / (ua0 - p_min)), 1.0);
          if (theta != 1.0) {
#pragma unroll
            for (int a = 0; a < M; ++a)
#pragma unroll
              for (int b = 0; b < M; ++b)
                if (a != 0 || b != 0) acc[a][b] = theta * acc[a][b];
          }
        }
        double t_min = 1.;
#pragma unroll 1
        for (int fam = 0; fam < 2; ++fam) {
          if (!ok) {
#pragma unroll 1
            for (int q = 0; q < M; ++q)
#pragma unroll 1
              for (int r = 0; r < GLL; ++r)
                PT[((q * GLL + r) * 4 + v) * 32 + lane] = eval_at<M>(acc, fam == 0 ? B.Pg[r] : B.P[q], fam == 0 ? B.P[q] : B.Pg[r]);
          }
          __syncthreads();
          if (!ok) {
#pragma unroll 1
            for (int k = v; k < M * GLL; k += 4) {
              const double* p = PT + (k * 4) * 32 + lane;
              const double pt[4] = {p[0], p[32], p[64], p[96]};
              const double ua[4] = {ua0, ua1, ua2, ua3};
              const fastm::Prim w = fastm::prim(P, pt[0], pt[1], pt[2], pt[3]);
              double t = 1.;
              if (!(w.p > P.eps)) t = solve_for_t(P, pt, ua);
              if (t_min >= t) t_min = t;
            }
          }
          __syncthreads();
        }
        LIM[v * 32 + lane] = t_min;
        __syncthreads();
        t_min = fmin(fmin(LIM[lane], LIM[32 + lane]), fmin(LIM[64 + lane], LIM[96 + lane]));
        if (!ok && t_min != 1.0) {
#pragma unroll
          for (int a = 0; a < M; ++a)
#pragma unroll
            for (int b = 0; b < M; ++b)
              if (a != 0 || b != 0) acc[a][b] = t_min * acc[a][b];
        }
        __syncthreads();                                     // the next row's nodal values go where PT is
      }
    } else {
      __syncthreads();                                       // phase A of the next row overwrites what phase C has read
    }
    if (OUT2) {
#pragma unroll
      for (int b = 0; b < M; ++b)
#pragma unroll
        for (int a = 0; a < M; ++a) C.out2[pe + (size_t)(b * M + a) * g.ne] = fma(C.k3, acc[a][b], W[OUT2 ? a : 0][OUT2 ? b : 0]);
    }
#pragma unroll
    for (int b = 0; b < M; ++b)
#pragma unroll
      for (int a = 0; a < M; ++a) out[pe + (size_t)(b * M + a) * g.ne] = acc[a][b];
    // slab boundary-row launches only (StageCoef): the same numbers straight into the neighbour's ghost row -- peer memory over
    // NVLink.  Kept behind one uniform test and after the ordinary stores so that the other launches pay nothing for it; the
    // second result is read back from where this thread has just put it.
    if (C.peer_lo != nullptr || C.peer_hi != nullptr) {
      double* pp = j == 1 ? C.peer_lo : j == g.ny - 2 ? C.peer_hi : nullptr;
      const size_t pne = j == 1 ? C.peer_lo_ne : C.peer_hi_ne;
      if (pp) {
#pragma unroll
        for (int b = 0; b < M; ++b)
#pragma unroll
          for (int a = 0; a < M; ++a) pp[((size_t)v * NM + b * M + a) * pne + ic0 + lane] = acc[a][b];
      }
      if (OUT2) {
        double* pp2 = j == 1 ? C.peer2_lo : j == g.ny - 2 ? C.peer2_hi : nullptr;
        if (pp2) {
#pragma unroll 1
          for (int m = 0; m < NM; ++m) pp2[((size_t)v * NM + m) * pne + ic0 + lane] = C.out2[pe + (size_t)m * g.ne];
        }
      }
    }
  }
}

}}  // namespace wb::dg
