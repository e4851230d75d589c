// TMA-fed variant of the fused 2D FV RK-stage kernel (production path on sm_100a).
//
// Same march as k_stage_march (a warp owns 31 output columns and walks up a strip of rows; every face flux is
// evaluated once; identical arithmetic, hence bit-identical results), but the rows no longer travel through
// per-lane global loads and register prefetch buffers:
//   * lane 0 of every warp issues ONE 3-D TMA tensor load per input array and row
//     (box = 34 columns x 1 row x the 4 planes of the delta-form state, covering the warp's 32 columns plus the left halo) into a per-warp ring of shared-memory row slots, completion signalled on one mbarrier per slot;
//   * the box may start anywhere whose BYTE offset is a multiple of 16 (measured on B200: an odd FP64 column
//     coordinate raises "illegal instruction"; negative even ones are fine), so it starts at the even column
//     x0 <= c0-1, and out-of-range columns (< 0, >= nx) are zero-filled by the hardware: no per-lane address
//     arithmetic, no index clamping and no lane-0 halo load;
//   * all lanes read their own column (row j+1) and their left neighbour (row j) from the slot with LDS;
//     the left-neighbour shuffle and its lane-0 select disappear;
//   * a slot is re-armed for row p+DEPTH as soon as row p is finished: DEPTH-2 rows are always in flight
//     without holding registers (the LDG version stalled 40 % of its time on the first use of a prefetched row).
// Warps stay independent: no block barrier anywhere -- a CTA is ONE warp (TMA_WARPS), so that every TMA operand follows from
// blockIdx and is provably uniform.  The main loop takes four rows per trip: the ring slot of a row (q & 3) is then a
// compile-time constant and every slot / barrier address is base + immediate; the strip's y-table entries are staged once in
// shared memory.  profiles/sass_loop.py counts the loop: 180 FP64 + 126.5 other instructions per row in stage 1.
// In a slab's boundary-row launch (one-row strips, the generic "tail" instantiation) the results are also stored into the
// neighbours' ghost rows: peer memory, StageArgs::peer_*.
#pragma once
#include <cuda.h>
#include <cstdint>

namespace wb { namespace fv2d {

constexpr int TMA_BOXW = 34;                          // box columns (272 B: multiple of 16 B as TMA requires)
constexpr int TMA_PLANE_B = TMA_BOXW * 8;             // bytes per plane row in a slot
constexpr int TMA_IN_B = 4 * TMA_PLANE_B;             // 1088: one row of a 4-plane field (delta form)
constexpr int TMA_IN_PAD = 1152;                      // slots start on 128-byte boundaries
constexpr int TMA_SLOT_B = TMA_IN_PAD;
constexpr int TMA_DEPTH = 4;                          // state ring: rows p, p+1 in use, p+2, p+3 in flight
#ifndef TMA_BDEPTH_N
#define TMA_BDEPTH_N 4
#endif
constexpr int TMA_BDEPTH = TMA_BDEPTH_N;             // u^n ring (stage 2): row q in use, the others in flight
constexpr int TMA_BARS_B = 128;
constexpr int TMA_MAX_ROWS = 64;                      // rows per strip the per-warp y tables are sized for
constexpr int TMA_TAB_N = TMA_MAX_ROWS + 4;           // entries per table
constexpr int TMA_TAB_B = 2 * TMA_TAB_N * 8;          // exp(-a yf), exp(-a yc) of the strip's rows (read with LDS, one per row)
// Warps per CTA.  The warps never talk to each other, so a CTA is ONE warp: its column block then follows from blockIdx
// alone, every TMA operand (box coordinates, slot and barrier addresses) is provably warp-uniform, and ptxas issues
// UTMALDG straight from uniform registers instead of wrapping it in an ELECT / R2UR / BRA.U.ANY uniformisation loop.
#ifndef TMA_WARPS_N
#define TMA_WARPS_N 1
#endif
constexpr int TMA_WARPS = TMA_WARPS_N;
__host__ __device__ constexpr int tma_warp_bytes(int mode) {
  return TMA_DEPTH * TMA_SLOT_B + (mode == 2 ? TMA_BDEPTH * TMA_IN_PAD : 0) + TMA_BARS_B + TMA_TAB_B;
}

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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void st_if(double* p, double v, bool on) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n@q st.global.f64 [%0], %1;\n}" ::"l"(p), "d"(v), "r"((int)on) : "memory");
}

struct TmaCtx {
  int lane, i, jb, np, nrows, jmin, jmax, x0, own, jgoff;
  unsigned jgspan;              // row j is interior  <=>  (unsigned)(j + jgoff) < jgspan  (0 < g.j0 + j < ny - 1)      // x0: first box column (even); own: this lane's column in the box
  bool writer, col_interior;
  double exf_i, exc_i, dt;
  const unsigned char* wsm;     // this warp's shared-memory region (generic pointer)
  uint32_t ring, bring, bars;   // shared-space addresses: state ring, u^n ring, mbarriers
  const double* tab;            // y tables of the strip: tab[q] = exp(-a yf(jb+q+1)), tab[TMA_TAB_N + q] = exp(-a yc(jb+q))
};

static_assert(TMA_DEPTH == 4 && TMA_BDEPTH == 4, "the slot arithmetic of the unrolled loop assumes rings of four");

// lane 0: arm slot p & (DEPTH-1) and ask TMA for local row jb-1+p (clamped like the LDG kernel: ghost rows exist
// only where a neighbouring slab does)
template <int MODE>
__device__ __forceinline__ void tma_issue_row(const TmaCtx& c, const CUtensorMap* m_in, int p) {
  const int s = p & (TMA_DEPTH - 1);
  const int row = max(c.jmin, min(c.jb - 1 + p, c.jmax)) + 1;
  const uint32_t bar = c.bars + 8u * s, dst = c.ring + (uint32_t)(s * TMA_SLOT_B);
  mbar_expect_tx(bar, TMA_IN_B);
  tma_load_3d(dst, m_in, c.x0, row, 0, bar);
}
__device__ __forceinline__ void tma_issue_base(const TmaCtx& c, const CUtensorMap* m_base, int q) {
  const int s = q & (TMA_BDEPTH - 1);
  const uint32_t bar = c.bars + 8u * (TMA_DEPTH + s), dst = c.bring + (uint32_t)(s * TMA_IN_PAD);
  mbar_expect_tx(bar, TMA_IN_B);
  tma_load_3d(dst, m_base, c.x0, c.jb + q + 1, 0, bar);
}
// The main loop is unrolled four rows deep, so a row knows S = q & 3 at compile time and every slot below is
// `base + immediate`; S = -1 (the ragged tail of a strip, the cold exact path) computes the slots from q.
template <int S> __device__ __forceinline__ int ring_slot(int q, int ahead) {
  return (S >= 0) ? ((S + ahead) & 3) : ((q + ahead) & 3);
}
// Re-arm for row q: ring row q (slot q & 3, dead since the previous row) takes ring row q + DEPTH, the slot of u^n row
// q - 1 takes u^n row q + 3.  No lower clamp here: jb - 1 + p >= 3 > jmin for p >= DEPTH.
// Everything but the lane test is computed by the whole warp (row, arm): values the compiler can see are warp-uniform live
// in uniform registers, and UTMALDG then needs no uniformisation loop.
template <int MODE, int S>
__device__ __forceinline__ void tma_rearm(const TmaCtx& c, const CUtensorMap* m_in, const CUtensorMap* m_base, int q) {
  const int over = c.jb + q + TMA_DEPTH - 1 - c.jmax;       // min() without a vector min: shifts and masks exist on the
  int neg;                                                   // uniform datapath, so the row stays in a uniform register
  asm("{\n.reg .s32 t;\nshr.s32 t, %1, 31;\nand.b32 %0, %1, t;\n}" : "=r"(neg) : "r"(over));      // (asm: NVVM turns the C++ form back into min)
  const int row = c.jmax + neg + 1;
  const bool arm = (q + TMA_DEPTH < c.np), arm_b = (MODE == 2) && (q >= 1) && (q + 3 < c.nrows);
  if (c.lane == 0 && arm) {
    const int s = ring_slot<S>(q, 0);
    const uint32_t bar = c.bars + 8u * s, dst = c.ring + (uint32_t)(s * TMA_SLOT_B);
    mbar_expect_tx(bar, TMA_IN_B);
    tma_load_3d(dst, m_in, c.x0, row, 0, bar);
  }
  if (MODE == 2 && c.lane == 0 && arm_b) {
    const int s = ring_slot<S>(q, 3);
    const uint32_t bar = c.bars + 8u * (TMA_DEPTH + s), dst = c.bring + (uint32_t)(s * TMA_IN_PAD);
    mbar_expect_tx(bar, TMA_IN_B);
    tma_load_3d(dst, m_base, c.x0, c.jb + q + 4, 0, bar);
  }
}
// box column `col` (c.own = this lane's cell, c.own - 1 = its left neighbour) of the ring slot s
__device__ __forceinline__ Cell tma_read_slot(const TmaCtx& c, int s, int col) {
  const unsigned char* sp = c.wsm + col * 8 + s * TMA_SLOT_B;
  Cell r;
  r.d0 = *reinterpret_cast<const double*>(sp);
  r.d1 = *reinterpret_cast<const double*>(sp + TMA_PLANE_B);
  r.d2 = *reinterpret_cast<const double*>(sp + 2 * TMA_PLANE_B);
  r.d3 = *reinterpret_cast<const double*>(sp + 3 * TMA_PLANE_B);
  return r;
}

// Arithmetic of one row (strip-relative index q, ring row p = q+1, local row j = jb+q): reads row j+1 (own column),
// row j (left neighbour) and u^n of row j from the ring; (cur,Gb) in; nxt, Gt and the new cell values out.
struct RowOut { Cell nxt; FaceFlux Gt; double n0, n1, n2, n3; };
template <int MODE, bool EXACT, int S>
__device__ __forceinline__ RowOut tma_row_math(const StageArgs& A, const Grid& g, const Phys& P, const TmaCtx& c, int q,
                                               const Cell& cur, const FaceFlux& Gb, double ey, double ex, bool& ok) {
  const int j = c.jb + q;
  RowOut o;
  o.nxt = tma_read_slot(c, ring_slot<S>(q, 2), c.own);
  const Cell lft = tma_read_slot(c, ring_slot<S>(q, 1), c.own - 1);
  double b0 = 0, b1 = 0, b2 = 0, b3 = 0;
  if (MODE == 2) {
    const unsigned char* sp = c.wsm + c.own * 8 + (TMA_DEPTH * TMA_SLOT_B + ring_slot<S>(q, 0) * TMA_IN_PAD);
    b0 = *reinterpret_cast<const double*>(sp);
    b1 = *reinterpret_cast<const double*>(sp + TMA_PLANE_B);
    b2 = *reinterpret_cast<const double*>(sp + 2 * TMA_PLANE_B);
    b3 = *reinterpret_cast<const double*>(sp + 3 * TMA_PLANE_B);
  }
  // ---- top y-face (j+1; normal = y) and left x-face (i; normal = x): four states in lock-step
  FaceFlux Fl;
  {
    const FaceIn fy = {P.rho0 * ey, P.pe1 * ey, cur.d0, cur.d2, cur.d1, cur.d3, o.nxt.d0, o.nxt.d2, o.nxt.d1, o.nxt.d3};
    const FaceIn fx = {P.rho0 * ex, P.pe1 * ex, lft.d0, lft.d1, lft.d2, lft.d3, cur.d0, cur.d1, cur.d2, cur.d3};
    faces_llf2<EXACT>(P, fy, fx, o.Gt, Fl, ok);
  }
  // ---- right x-face (i+1) from lane+1
  FaceFlux Fr;
  Fr.f0 = __shfl_down_sync(0xffffffffu, Fl.f0, 1); Fr.fn = __shfl_down_sync(0xffffffffu, Fl.fn, 1);
  Fr.ft = __shfl_down_sync(0xffffffffu, Fl.ft, 1); Fr.f3 = __shfl_down_sync(0xffffffffu, Fl.f3, 1);
  // ---- dudt in the reference's order (benchmark_2d.f90:601-607), RK axpy
  const bool interior = c.col_interior && ((unsigned)(j + c.jgoff) < c.jgspan);
  cell_update<MODE>(P, cur, Fl, Fr, Gb, o.Gt, interior, c.dt, b0, b1, b2, b3, o.n0, o.n1, o.n2, o.n3);
  return o;
}
// cold path, out of line and with by-value arguments only (taking addresses of the kernel's structs would move them
// to local memory for the whole kernel): some state of the warp's row sits at a floor of the sound-speed formula
template <int MODE>
__device__ __noinline__ RowOut tma_row_exact(StageArgs A, Grid g, Phys P, TmaCtx c, int q, Cell cur, FaceFlux Gb, double ey,
                                             double ex) {
  bool ok = true;
  return tma_row_math<MODE, true, -1>(A, g, P, c, q, cur, Gb, ey, ex, ok);
}

// One row including the ring bookkeeping and the stores.  (cur,Gb) in, (nxt,Gt) out as in march_row.  ph = (q >> 2) & 1.
template <int MODE, int S>
__device__ __forceinline__ void tma_row(const StageArgs& A, const Grid& g, const Phys& P, const TmaCtx& c,
                                        const CUtensorMap* m_in, const CUtensorMap* m_base, int q, uint32_t ph,
                                        const Cell& cur, Cell& nxt, const FaceFlux& Gb, FaceFlux& Gt, double& spd) {
  const int j = c.jb + q;
  // ---- the previous row left ring row q and u^n row q-1 dead: re-arm their slots (kept next to the barrier wait
  //      so that the arithmetic of a row stays one basic block for the instruction scheduler)
  __syncwarp();
  tma_rearm<MODE, S>(c, m_in, m_base, q);
  // ---- y tables of this row
  const double tyf = c.tab[q], tyc = c.tab[TMA_TAB_N + q];
  const double ey = c.exc_i * tyf, ex = c.exf_i * tyc, ec = c.exc_i * tyc;
  // ---- ring row q+2 (local row j+1, own column) must have landed; row q+1 (left neighbour) landed a row ago
  {
    const int s = ring_slot<S>(q, 2);
    const uint32_t par = (S >= 0) ? ((S + 2 >= 4) ? (ph ^ 1u) : ph) : (uint32_t)(((q + 2) >> 2) & 1);
    mbar_wait(c.bars + 8u * s, par);
  }
  if (MODE == 2) mbar_wait(c.bars + 8u * (TMA_DEPTH + ring_slot<S>(q, 0)), ph);
  bool ok = true;
  RowOut o = tma_row_math<MODE, false, S>(A, g, P, c, q, cur, Gb, ey, ex, ok);
  if (!__all_sync(0xffffffffu, ok)) o = tma_row_exact<MODE>(A, g, P, c, q, cur, Gb, ey, ex);
  nxt = o.nxt;
  Gt = o.Gt;
  // predicated stores (no branch: lane 31 and the lanes beyond nx simply do not write)
  double* dst = A.out + ((size_t)(j + 1) * g.pitch + c.i);
  st_if(dst, o.n0, c.writer); st_if(dst + g.plane, o.n1, c.writer); st_if(dst + 2 * g.plane, o.n2, c.writer);
  st_if(dst + 3 * g.plane, o.n3, c.writer);
  if (S < 0) {      // only the generic (tail) instantiation: the slab's boundary rows are one-row strips
    double* pr = nullptr;
    size_t pp = 0;
    if (j == 0 && A.peer_lo) { pr = A.peer_lo; pp = A.peer_lo_plane; }
    else if (j == g.nyl - 1 && A.peer_hi) { pr = A.peer_hi; pp = A.peer_hi_plane; }
    if (pr) {       // the same row into the neighbour's ghost row (peer memory, NVLink)
      pr += c.i;
      st_if(pr, o.n0, c.writer); st_if(pr + pp, o.n1, c.writer); st_if(pr + 2 * pp, o.n2, c.writer); st_if(pr + 3 * pp, o.n3, c.writer);
    }
  }
  if (MODE == 2) {
    const double s = centre_speed(A, g, P, (size_t)(j + 1) * g.pitch + min(c.i, g.nx - 1), ec, o.n0, o.n1, o.n2, o.n3);
    spd = (c.writer && s > spd) ? s : spd;
  }
}

template <int MODE, int MB>
__global__ void __launch_bounds__(TMA_WARPS * 32, MB * MARCH_WARPS / TMA_WARPS)
k_stage_tma(const __grid_constant__ CUtensorMap m_in, const __grid_constant__ CUtensorMap m_base, StageArgs A, Grid g, Phys P,
            int R) {
  extern __shared__ __align__(128) unsigned char tma_smem[];
  TmaCtx c;
  c.dt = 0.0;
  if (MODE != 0) {
    if (step_done(A.ctrl, A.parity, A.tend, A.max_iter)) {
      if (MODE == 2 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && A.row_begin == 0)
        carry_forward(A.ctrl, A.parity);
      return;
    }
    c.dt = step_dt(A.ctrl, A.parity, P);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && A.row_begin == 0)
      bookkeeping<MODE>(A.ctrl, A.parity, c.dt);
  }
  c.lane = threadIdx.x & 31;
  const int warp = (TMA_WARPS == 1) ? 0 : __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int c0 = (blockIdx.x * TMA_WARPS + warp) * MARCH_OUT;
  if (c0 >= g.nx) return;                                 // whole warp: no block barriers in this kernel
  c.jb = A.row_begin + blockIdx.y * R;
  const int je = min(c.jb + min(R, A.rows_cap), A.row_end);
  if (c.jb >= je) return;
  c.nrows = je - c.jb;
  c.np = c.nrows + 2;                                       // ring rows p = 0..np-1  <->  local rows jb-1 .. je
  c.i = c0 + c.lane;
  c.x0 = (c0 - 1) & ~1;                                     // even: 16-byte aligned box start (also for c0 = 0: -2)
  c.own = c.i - c.x0;
  c.jmin = (g.j0 > 0) ? -1 : 0;
  c.jmax = (g.j0 + g.nyl < g.ny) ? g.nyl : g.nyl - 1;
  c.jgoff = g.j0 - 1;
  c.jgspan = (unsigned)max(g.ny - 2, 0);
  const int ic = min(c.i, g.nx - 1);
  c.exf_i = A.exf[ic];
  c.exc_i = A.exc[ic];
  c.writer = (c.lane < MARCH_OUT) && (c.i < g.nx);
  c.col_interior = (c.i > 0) && (c.i < g.nx - 1);
  c.wsm = tma_smem + warp * tma_warp_bytes(MODE);
  c.ring = smem_u32(c.wsm);
  c.bring = c.ring + TMA_DEPTH * TMA_SLOT_B;
  c.bars = c.bring + (MODE == 2 ? TMA_BDEPTH * TMA_IN_PAD : 0);
  double* tabw = reinterpret_cast<double*>(const_cast<unsigned char*>(c.wsm) + (TMA_DEPTH * TMA_SLOT_B +
                                           (MODE == 2 ? TMA_BDEPTH * TMA_IN_PAD : 0) + TMA_BARS_B));
  c.tab = tabw;
  // ---- y tables of the strip: row q multiplies its face / centre x-table entries by exp(-a yf(jb+q+1)), exp(-a yc(jb+q))
  for (int e = c.lane; e < TMA_TAB_N; e += 32) {
    tabw[e] = A.eyf[min(c.jb + e + 1, g.nyl)];
    tabw[TMA_TAB_N + e] = A.eyc[min(c.jb + e, g.nyl - 1)];
  }

  // ---- barriers, then the first DEPTH rows
  if (c.lane == 0) {
#pragma unroll
    for (int s = 0; s < TMA_DEPTH + (MODE == 2 ? TMA_BDEPTH : 0); ++s) mbar_init(c.bars + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
    for (int p = 0; p < TMA_DEPTH; ++p)
      if (p < c.np) tma_issue_row<MODE>(c, &m_in, p);
    if (MODE == 2) {
#pragma unroll
      for (int q = 0; q < TMA_BDEPTH; ++q)
        if (q < c.nrows) tma_issue_base(c, &m_base, q);
    }
  }
  __syncwarp();

  // ---- prologue: bottom face of the strip from rows jb-1 (p = 0) and jb (p = 1)
  Cell ca, cb;
  FaceFlux Ga, Gb2;
  {
    mbar_wait(c.bars, 0);
    const Cell bel = tma_read_slot(c, 0, c.own);
    mbar_wait(c.bars + 8u, 0);
    ca = tma_read_slot(c, 1, c.own);
    const double e = c.exc_i * A.eyf[c.jb];
    Ga = face_llf(P, P.rho0 * e, P.pe1 * e, bel.d0, bel.d2, bel.d1, bel.d3, ca.d0, ca.d2, ca.d1, ca.d3);
  }
  double spd = 0.0;
  int q = 0;
  uint32_t ph = 0;
#pragma unroll 1
  for (; q + 3 < c.nrows; q += 4) {     // four rows per trip: compile-time ring slots, (ca,Ga)->(cb,Gb2)->(ca,Ga) without moves
    tma_row<MODE, 0>(A, g, P, c, &m_in, &m_base, q, ph, ca, cb, Ga, Gb2, spd);
    tma_row<MODE, 1>(A, g, P, c, &m_in, &m_base, q + 1, ph, cb, ca, Gb2, Ga, spd);
    tma_row<MODE, 2>(A, g, P, c, &m_in, &m_base, q + 2, ph, ca, cb, Ga, Gb2, spd);
    tma_row<MODE, 3>(A, g, P, c, &m_in, &m_base, q + 3, ph, cb, ca, Gb2, Ga, spd);
    ph ^= 1u;
  }
#pragma unroll 1
  for (; q < c.nrows; ++q) {            // ragged tail of the last strip (cold: strips are 32 rows)
    tma_row<MODE, -1>(A, g, P, c, &m_in, &m_base, q, (uint32_t)((q >> 2) & 1), ca, cb, Ga, Gb2, spd);
    ca = cb;
    Ga = Gb2;
  }
  if (MODE == 2) {
    spd = warp_max(spd);
    if (c.lane == 0) atomic_max_nonneg(&A.ctrl->cmax_bits[A.parity ^ 1], spd);
  }
}

}}  // namespace wb::fv2d
