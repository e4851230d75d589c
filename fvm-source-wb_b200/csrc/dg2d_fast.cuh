// Fused DG 2D RK-stage arithmetic ("arith 0") and its one-thread-per-element kernel k_dg_stage_fast (any grid; the
// production kernel for nx % 32 == 0 is k_dg_stage_split, dg2d_split.cuh, which runs the same arithmetic with the element
// split over four threads): compute_update + RK combination + element-local positivity limiter
// ('ONP') in ONE launch per stage.  Same formulas as the reference-order kernels of dg2d.cu, but
//   * every tensor-product contraction is sum-factorised (2 x M^3 instead of M^4 multiply-adds per variable),
//     with the quadrature weights folded into the basis tables (Pw = P*w, dPw = P'*w);
//   * divisions / square roots use the Newton-refined MUFU seeds of the FV kernels, products are fused (FMA);
//   * dudt never touches memory: the stage output  c0*A0 + c1*A1 + c2*A2 + c3*A3 + cd*dt*L(in)  is formed in
//     registers, limited, and written once.
// Results differ from the reference order by a few ulp per operation (parity bar 1e-12, tests/test_dg2d_gpu.py).
// Included by dg2d.cu (uses its DgGrid / DgPhys / DgCtrl / Basis definitions).
#pragma once

#ifndef DG_ONP_SLOW_ATTR
#define DG_ONP_SLOW_ATTR __forceinline__   /* measured: __noinline__ puts the accumulators in local memory, 2.6e9 -> 1.65e9 */
#endif

namespace wb { namespace dg {

struct FastBasis {
  double P[MAXM][MAXM];     // P[q][n]
  double Pw[MAXM][MAXM];    // P[q][n] * w[q]
  double dPw[MAXM][MAXM];   // P'[q][n] * w[q]
  double Em[MAXM], Ep[MAXM];
  double Pg[MAXM][MAXM];    // legendre(x_gll(r), n)
  double Pwh[MAXM][MAXM];   // 0.5 * Pw (exact): edge integrals of doubled fluxes, see fastm::llf
  double EpEp[MAXM][MAXM];  // Ep[a] * Ep[b]: bound of |P_a P_b| on the element ('ONP' sufficient test)
  int gll;
};

struct StageCoef {          // out = limiter( c0*A0 + c1*A1 + cd*dt*L(in) )
  const double *A0, *A1;
  double c0, c1, cd;
  int na;                   // number of A terms (1 or 2)
  // optional second result of the same launch (SSPRK(5,4): w5 needs L(w3) again, :700-704):
  //   out2 = k0*B0 + k1*B1 + k2*in + k3*out + ke*dt*L(in)
  double* out2;
  const double *B0, *B1;
  double k0, k1, k2, k3, ke;
  // slab boundary-row launches with peer-memory ghost rows (k_dg_stage_split only): local row 1 is ALSO stored into the ghost
  // row `peer_lo` points at (plane (0,0), column 0 of the lower neighbour's top ghost row -- or of the rank's own bottom ghost
  // row at a clamped global edge), local row ny-2 into `peer_hi`; plane strides of those fields in peer_*_ne; peer2_* = the
  // same for out2.  All null in every other launch.
  double *peer_lo, *peer_hi, *peer2_lo, *peer2_hi;
  size_t peer_lo_ne, peer_hi_ne;
};

namespace fastm {
__device__ __forceinline__ double rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  double t = fma(e, e, e);
  return fma(y, t, y);
}
__device__ __forceinline__ double sqrt_pos(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double t = a * y;
  double h = fma(-t, y, 1.0);
  double p = fma(0.375, h, 0.5);
  p = p * h;
  return fma(t, p, t);
}
// primitive variables with the reference's density floor (2d/benchmark_2d_dg.f90:891-902)
struct Prim { double w0, vx, vy, p, r; };
__device__ __forceinline__ Prim prim(const DgPhys& P, double rho, double mx, double my, double E) {
  Prim w;
  w.w0 = rho > P.rho_floor ? rho : P.rho_floor;      // fmax(rho, 10e-10 real(4)) for every rho, three instructions instead of five
  w.r = rcp(w.w0);
  w.vx = mx * w.r;
  w.vy = my * w.r;
  w.p = P.gm1a * fma(-0.5 * w.w0, fma(w.vy, w.vy, w.vx * w.vx), E);
  return w;
}
// 'hll2' / 'hllc' (compute_hllflux :1008-1026, compute_hllcflux :1030-1134): the reference-order routines of dg2d.cu, out of
// line and with by-value arguments so that the fused kernel's registers and instruction stream do not pay for them
struct Flux4 { double f[4]; };
template <int DIR>
__device__ __noinline__ Flux4 other_flux(DgPhys P, double a0, double a1, double a2, double a3, double b0, double b1, double b2,
                                         double b3) {
  const double ul[4] = {a0, a1, a2, a3}, ur[4] = {b0, b1, b2, b3};
  Flux4 r;
  num_flux<DIR>(P, ul, ur, r.f);
  return r;
}
// local Lax-Friedrichs at one face point (compute_llflux :968-988 with compute_flux_int :946-965), DIR 1 = x, 2 = y
// ANYFLUX = false: kernel instantiation for flux_id 0 | 1 only (no call in the instruction stream of the hot path)
#ifndef DG_LLF_ATTR
#define DG_LLF_ATTR __forceinline__
#endif
// DOUBLED: returns 2*flux (exactly: the two halvings of 0.5*(fb+fa) + 0.5*cmax*(ul-ur) are powers of two) for callers that
// fold the 0.5 into their quadrature table -- five FP64 instructions less per face point, same bits after the edge integral
template <int DIR, bool ANYFLUX, bool DOUBLED = false>
__device__ DG_LLF_ATTR void llf(const DgPhys& P, const double ul[4], const double ur[4], double nf[4]) {
  if (P.flux_id != 1) {
    nf[0] = nf[1] = nf[2] = nf[3] = 0.0;
    if (ANYFLUX && P.flux_id != 0) {
      const Flux4 r = other_flux<DIR>(P, ul[0], ul[1], ul[2], ul[3], ur[0], ur[1], ur[2], ur[3]);
      const double sc = DOUBLED ? 2.0 : 1.0;
      nf[0] = sc * r.f[0]; nf[1] = sc * r.f[1]; nf[2] = sc * r.f[2]; nf[3] = sc * r.f[3];
    }
    return;
  }
  const Prim a = prim(P, ul[0], ul[1], ul[2], ul[3]);
  const Prim b = prim(P, ur[0], ur[1], ur[2], ur[3]);
  // w0 >= 1e-9 > 1e-10, so max(w0,1d-10) = w0 and its reciprocal is already known
  const double pa = a.p > P.p_floor ? a.p : P.p_floor, pb = b.p > P.p_floor ? b.p : P.p_floor;      // max(p, 1d-10)
  const double ca = sqrt_pos(P.gamma * pa * a.r), cb = sqrt_pos(P.gamma * pb * b.r);
  const double vna = (DIR == 1) ? a.vx : a.vy, vnb = (DIR == 1) ? b.vx : b.vy;
  const double sa = fabs(vna + ca), sb = fabs(vnb + cb);
  const double cm = sb > sa ? sb : sa;
  const double hc = DOUBLED ? cm : 0.5 * cm;
  double fa[4], fb[4];
  const double ta = a.w0 * a.vx * a.vy, tb = b.w0 * b.vx * b.vy;
  if (DIR == 1) {
    fa[0] = a.vx * ul[0]; fa[1] = fma(a.vx, ul[1], a.p); fa[2] = ta; fa[3] = a.vx * (ul[3] + a.p);
    fb[0] = b.vx * ur[0]; fb[1] = fma(b.vx, ur[1], b.p); fb[2] = tb; fb[3] = b.vx * (ur[3] + b.p);
  } else {
    fa[0] = a.vy * ul[0]; fa[1] = ta; fa[2] = fma(a.vy, ul[2], a.p); fa[3] = a.vy * (ul[3] + a.p);
    fb[0] = b.vy * ur[0]; fb[1] = tb; fb[2] = fma(b.vy, ur[2], b.p); fb[3] = b.vy * (ur[3] + b.p);
  }
#pragma unroll
  for (int v = 0; v < 4; ++v) nf[v] = DOUBLED ? fma(hc, ul[v] - ur[v], fb[v] + fa[v]) : fma(hc, ul[v] - ur[v], 0.5 * (fb[v] + fa[v]));
}
#ifndef DG_LLF_CALL
#define DG_LLF_CALL 1
#endif
#if DG_LLF_CALL
// out-of-line copy of the LLF point evaluation, all arguments by value (registers).  The stage kernel is ~8800 SASS
// instructions of straight-line code and ncu shows `no_instruction` at 1.4 warps per issue: the 12 inlined copies of this
// routine were a quarter of the code a warp streams through.  Measured: 2.62e9 -> 2.79e9 element-stages/s at 4096^2.
template <int DIR>
__device__ __noinline__ Flux4 llf_call(double gamma, double gm1a, double a0, double a1, double a2, double a3, double b0, double b1,
                                       double b2, double b3) {
  DgPhys P;
  P.gamma = gamma; P.gm1a = gm1a; P.flux_id = 1; P.rho_floor = (double)10e-10f; P.p_floor = 1e-10;
  const double ul[4] = {a0, a1, a2, a3}, ur[4] = {b0, b1, b2, b3};
  Flux4 r;
  llf<DIR, false>(P, ul, ur, r.f);
  return r;
}
#endif
}  // namespace fastm

// trace of one variable on one side: SIDE 0 left, 1 right (points along y), 2 bottom, 3 top (points along x)
// Structural constants of the basis (checked at create, dg2d.cu): P_0 = 1, so Em[0] = Ep[0] = P[q][0] = 1 and dPw[q][0] = 0
// exactly.  The fused kernels use them: a sum that starts with fma(x, 1, 0) starts with x, terms with a zero factor are
// dropped (equal results up to the sign of an exact zero).  For an odd number of Gauss points the middle node is the origin
// (the reference's Newton iterate is 0 or ~1e-17: snapped to 0 in the FUSED tables only), where the odd polynomials and the
// derivatives of the even ones vanish: zP / zD name those table entries, and their terms are dropped as well.
template <int M> __host__ __device__ constexpr bool zP(int q, int n) { return (M & 1) && q == M / 2 && (n & 1); }    // P, Pw, Pwh
template <int M> __host__ __device__ constexpr bool zD(int q, int n) { return (M & 1) && q == M / 2 && !(n & 1); }   // dPw
template <int M, int SIDE>
__device__ __forceinline__ void trace1(const double (&d)[M][M], const FastBasis& B, double (&out)[M]) {
  double t[M];
  if (SIDE < 2) {
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double a = d[0][j];
#pragma unroll
      for (int i = 1; i < M; ++i) a = fma(d[i][j], SIDE == 0 ? B.Em[i] : B.Ep[i], a);
      t[j] = a;
    }
  } else {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double a = d[i][0];
#pragma unroll
      for (int j = 1; j < M; ++j) a = fma(d[i][j], SIDE == 2 ? B.Em[j] : B.Ep[j], a);
      t[i] = a;
    }
  }
#pragma unroll
  for (int q = 0; q < M; ++q) {
    double a = t[0];
#pragma unroll
    for (int n = 1; n < M; ++n)
      if (!zP<M>(q, n)) a = fma(t[n], B.P[q][n], a);
    out[q] = a;
  }
}
template <int M>
__device__ __forceinline__ void load_var(const double* __restrict__ u, const DgGrid& g, int v, size_t e, double (&d)[M][M]) {
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int i = 0; i < M; ++i) d[i][j] = PL(u, g, v, j * M + i)[e];
}
// value of one variable at the point (P_x-row px[], P_y-row py[])
template <int M>
__device__ __forceinline__ double eval_at(const double (&d)[M][M], const double* px, const double* py) {
  double a = 0.0;
#pragma unroll
  for (int i = 0; i < M; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j) s = fma(d[i][j], py[j], s);
    a = fma(s, px[i], a);
  }
  return a;
}

// 'ONP' positivity limiter (2d/limiters.f90:478-654) on an element held in registers
template <int M>
__device__ DG_ONP_SLOW_ATTR void positivity_slow(const DgPhys& P, const FastBasis& B, double (&el)[4][M][M]);

template <int M>
__device__ __forceinline__ void positivity_fast(const DgPhys& P, const FastBasis& B, double (&el)[4][M][M]) {
  if (M == 1) return;
  const double ua[4] = {el[0][0][0], el[1][0][0], el[2][0][0], el[3][0][0]};
  // ---- cheap sufficient test.  |P_i| <= P_i(1) = sqrt(2i+1) on [-1,1], so every point value of variable v lies within
  //      R_v = sum_{(i,j) != (0,0)} |mode_ij| P_i(1) P_j(1) of its mean.  If the resulting lower bounds of the density and of
  //      the pressure stay above eps (with a margin far above the rounding of the point evaluations), every point of the
  //      set gives theta = 1 and t = 1 exactly and the limiter leaves the element untouched -- the common case on
  //      resolved data, ~60 instead of ~1400 FP64 instructions.
  {
    double R[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      double r = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j)
          if (i != 0 || j != 0) r = fma(fabs(el[v][i][j]), B.Ep[i] * B.Ep[j], r);
      R[v] = r;
    }
    const double rho_lo = ua[0] - R[0];
    const double mx_hi = fabs(ua[1]) + R[1], my_hi = fabs(ua[2]) + R[2], E_lo = ua[3] - R[3];
    const double margin = 1e-9 * (fabs(ua[0]) + R[0] + fabs(ua[3]) + R[3]);
    bool ok = false;
    if (rho_lo > P.eps + margin && rho_lo > (double)10e-10f) {
      const double p_lo = P.gm1a * (E_lo - 0.5 * (mx_hi * mx_hi + my_hi * my_hi) / rho_lo);
      ok = p_lo > P.eps + margin;
    }
    // the hint tells ptxas to move the ~1400 instructions of the slow path out of the instruction stream of the common case
    // (the kernel is short of instruction-fetch bandwidth: +1.6 %)
    if (__builtin_expect(ok, 1)) return;
  }
  positivity_slow<M>(P, B, el);
}

// the point evaluations of 'ONP' (elements that fail the sufficient test): out of line, so that the stage kernel's register
// allocation is not sized for them
template <int M>
__device__ DG_ONP_SLOW_ATTR void positivity_slow(const DgPhys& P, const FastBasis& B, double (&el)[4][M][M]) {
  const double ua[4] = {el[0][0][0], el[1][0][0], el[2][0][0], el[3][0][0]};
  double p_min = 1e300;
  for (int q = 0; q < M; ++q)
    for (int r = 0; r < B.gll; ++r) {
      p_min = fmin(p_min, eval_at<M>(el[0], B.Pg[r], B.P[q]));     // "left" family: GLL in x, GL in y
      p_min = fmin(p_min, eval_at<M>(el[0], B.P[q], B.Pg[r]));     // "right" family: GL in x, GLL in y
    }
  const double theta = fmin(fabs((ua[0] - P.eps) / (ua[0] - p_min)), 1.0);
  if (theta != 1.0) {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j)
        if (i != 0 || j != 0) el[0][i][j] = theta * el[0][i][j];
  }
  double t_min = 1.;
  for (int fam = 0; fam < 2; ++fam)
    for (int q = 0; q < M; ++q)
      for (int r = 0; r < B.gll; ++r) {
        const double* px = fam == 0 ? B.Pg[r] : B.P[q];
        const double* py = fam == 0 ? B.P[q] : B.Pg[r];
        double pt[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) pt[v] = eval_at<M>(el[v], px, py);
        const fastm::Prim w = fastm::prim(P, pt[0], pt[1], pt[2], pt[3]);
        double t = 1.;
        if (!(w.p > P.eps)) t = solve_for_t(P, pt, ua);      // rare: exact reference routine
        if (t_min >= t) t_min = t;
      }
  if (t_min != 1.0) {
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j)
          if (i != 0 || j != 0) el[v][i][j] = t_min * el[v][i][j];
  }
}

// One face of the element: traces of the element (SIDE_OWN) and of the neighbour across it (SIDE_NB, modes at eN),
// local Lax-Friedrichs at the M face points, and the edge integral accumulated into acc with its sign:
//   x faces: -/+ E[a] * sum_q F[q] Pw[q][b]        y faces: -/+ E[b] * sum_q G[q] Pw[q][a]
// FACE 0 left, 1 right, 2 bottom, 3 top.  Each face is evaluated by both adjacent elements (no inter-thread traffic).
// `src.template nb<FACE>(v, dn)` delivers the modes of variable v of the neighbour across the face (global memory or the
// TMA-staged shared-memory rows, see the two sources below).
// numerical flux at the M points of one face from the modes of the two elements that share it (low side first)
template <int M, int FACE, bool ANYFLUX>
__device__ __forceinline__ void face_flux_from_traces(const DgPhys& P, double (&to)[M][4], const double (&tn)[M][4]) {
#pragma unroll
  for (int q = 0; q < M; ++q) {
    double F[4];
    // low side first: (neighbour, own) on the left/bottom faces, (own, neighbour) on the right/top faces
#if DG_LLF_CALL
    if (!ANYFLUX && P.flux_id == 1) {
      const double* lo = (FACE == 0 || FACE == 2) ? tn[q] : to[q];
      const double* hi = (FACE == 0 || FACE == 2) ? to[q] : tn[q];
      const fastm::Flux4 r = fastm::llf_call<(FACE < 2) ? 1 : 2>(P.gamma, P.gm1a, lo[0], lo[1], lo[2], lo[3], hi[0], hi[1], hi[2], hi[3]);
      F[0] = r.f[0]; F[1] = r.f[1]; F[2] = r.f[2]; F[3] = r.f[3];
    } else
#endif
    {
    if (FACE == 0) fastm::llf<1, ANYFLUX>(P, tn[q], to[q], F);
    if (FACE == 1) fastm::llf<1, ANYFLUX>(P, to[q], tn[q], F);
    if (FACE == 2) fastm::llf<2, ANYFLUX>(P, tn[q], to[q], F);
    if (FACE == 3) fastm::llf<2, ANYFLUX>(P, to[q], tn[q], F);
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) to[q][v] = F[v];
  }
}
// edge integral of the face flux F[q][v] accumulated into acc with its sign (see face_term)
template <int M, int FACE>
__device__ __forceinline__ void face_accum(const FastBasis& B, const double (&F)[M][4], double (&acc)[4][M][M]) {
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    double s[M];
#pragma unroll
    for (int n = 0; n < M; ++n) {
      double a1 = 0.0;
#pragma unroll
      for (int q = 0; q < M; ++q)
        if (!zP<M>(q, n)) a1 = fma(F[q][v], B.Pw[q][n], a1);
      s[n] = a1;
    }
#pragma unroll
    for (int a = 0; a < M; ++a)
#pragma unroll
      for (int b = 0; b < M; ++b) {
        if (FACE == 0) acc[v][a][b] = fma(B.Em[a], s[b], acc[v][a][b]);     // + e2
        if (FACE == 1) acc[v][a][b] = fma(-B.Ep[a], s[b], acc[v][a][b]);    // - e1
        if (FACE == 2) acc[v][a][b] = fma(B.Em[b], s[a], acc[v][a][b]);     // + e4
        if (FACE == 3) acc[v][a][b] = fma(-B.Ep[b], s[a], acc[v][a][b]);    // - e3
      }
  }
}
template <int M, int FACE, bool ANYFLUX, class Src>
__device__ __forceinline__ void face_term(Src& src, const DgPhys& P, const FastBasis& B, const double (&d)[4][M][M],
                                          double (&acc)[4][M][M]) {
  constexpr int SIDE_OWN = FACE;                                   // own trace on that side
  constexpr int SIDE_NB = (FACE == 0) ? 1 : (FACE == 1) ? 0 : (FACE == 2) ? 3 : 2;   // neighbour's facing side
  double to[M][4], tn[M][4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    double t1[M], dn[M][M];
    trace1<M, SIDE_OWN>(d[v], B, t1);
#pragma unroll
    for (int q = 0; q < M; ++q) to[q][v] = t1[q];
    src.template nb<FACE>(v, dn);
    trace1<M, SIDE_NB>(dn, B, t1);
#pragma unroll
    for (int q = 0; q < M; ++q) tn[q][v] = t1[q];
  }
  face_flux_from_traces<M, FACE, ANYFLUX>(P, to, tn);
  face_accum<M, FACE>(B, to, acc);
}

// neighbour modes straight from global memory (L1/L2): the original data path
template <int M>
struct GlobalSrc {
  const double* __restrict__ in;
  const DgGrid& g;
  size_t e, eN[4];
  __device__ __forceinline__ void own(int v, double (&d)[M][M]) const { load_var<M>(in, g, v, e, d); }
  template <int FACE>
  __device__ __forceinline__ void nb(int v, double (&d)[M][M]) const { load_var<M>(in, g, v, eN[FACE], d); }
  __device__ __forceinline__ void x_faces_done() const {}
  __device__ __forceinline__ void bottom_face_done(const StageCoef&) const {}
  __device__ __forceinline__ void top_face_done(const StageCoef&) const {}
  // operands of the RK combination: A0, A1 and (for the second result) the stage input itself, mode m of variable v
  __device__ __forceinline__ double rk_a0(const StageCoef& C, int v, int m) const { return PL(C.A0, g, v, m)[e]; }
  __device__ __forceinline__ double rk_a1(const StageCoef& C, int v, int m) const { return PL(C.A1, g, v, m)[e]; }
  __device__ __forceinline__ double rk_in(int v, int m) const { return PL(in, g, v, m)[e]; }
};

template <int M, class Src, class Modes>
__device__ __forceinline__ void dg_stage_rest(Src& src, const Modes& d, double (&acc)[4][M][M], const StageCoef& C,
                                              double* __restrict__ out, const double* __restrict__ gx,
                                              const double* __restrict__ gy, const unsigned char* __restrict__ fz, const DgGrid& g,
                                              const DgPhys& P, const FastBasis& B, const DgCtrl* __restrict__ ctrl, int apply_onp,
                                              size_t e);

// the element's own modes held in registers (k_dg_stage_fast)
template <int M>
struct RegModes {
  const double (&d)[4][M][M];
  __device__ __forceinline__ void get(int v, double (&o)[M][M]) const {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) o[i][j] = d[v][i][j];
  }
};

// One element of one RK stage (everything but where the modes come from).
template <int M, bool ANYFLUX, class Src>
__device__ __forceinline__ void dg_stage_body(Src& src, const double* __restrict__ in, const StageCoef& C, double* __restrict__ out,
                                              const double* __restrict__ gx, const double* __restrict__ gy,
                                              const unsigned char* __restrict__ fz, const DgGrid& g, const DgPhys& P,
                                              const FastBasis& B, const DgCtrl* __restrict__ ctrl, int apply_onp, size_t e) {
  double acc[4][M][M];                     // -(e1-e2) - (e3-e4) + vol1 + vol2, then dudt, then the stage result
  double d[4][M][M];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    src.own(v, d[v]);
#pragma unroll
    for (int a = 0; a < M; ++a)
#pragma unroll
      for (int b = 0; b < M; ++b) acc[v][a][b] = 0.0;
  }
  // ---- faces first (they need the neighbours' modes; nothing of them stays live afterwards)
  face_term<M, 0, ANYFLUX>(src, P, B, d, acc);
  face_term<M, 1, ANYFLUX>(src, P, B, d, acc);
  src.x_faces_done();
  face_term<M, 2, ANYFLUX>(src, P, B, d, acc);
  src.bottom_face_done(C);
  face_term<M, 3, ANYFLUX>(src, P, B, d, acc);
  src.top_face_done(C);
  dg_stage_rest<M>(src, RegModes<M>{d}, acc, C, out, gx, gy, fz, g, P, B, ctrl, apply_onp, e);
}

// Everything of a stage after the face terms: nodal values, volume and source integrals, RK combination, 'ONP', stores.
// d = the element's modes (dead after the nodal evaluation), acc = the four edge integrals with their signs.
template <int M, class Src, class Modes>
__device__ __forceinline__ void dg_stage_rest(Src& src, const Modes& d, double (&acc)[4][M][M], const StageCoef& C,
                                              double* __restrict__ out, const double* __restrict__ gx,
                                              const double* __restrict__ gy, const unsigned char* __restrict__ fz, const DgGrid& g,
                                              const DgPhys& P, const FastBasis& B, const DgCtrl* __restrict__ ctrl, int apply_onp,
                                              size_t e) {
  double U[4][M][M];                       // nodal values -> nodal source -> out2 partial
  {
    // ---- nodal values (sum-factorised)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      double dv[M][M];
      d.get(v, dv);
#pragma unroll
      for (int qx = 0; qx < M; ++qx) {
        double a[M];
#pragma unroll
        for (int j = 0; j < M; ++j) {
          double s = dv[0][j];
#pragma unroll
          for (int i = 1; i < M; ++i)
            if (!zP<M>(qx, i)) s = fma(dv[i][j], B.P[qx][i], s);
          a[j] = s;
        }
#pragma unroll
        for (int qy = 0; qy < M; ++qy) {
          double s = a[0];
#pragma unroll
          for (int j = 1; j < M; ++j)
            if (!zP<M>(qy, j)) s = fma(a[j], B.P[qy][j], s);
          U[v][qx][qy] = s;
        }
      }
    }
  }

  // ---- volume terms: fluxes (and source) at the nodes, then vol1 + vol2 (+ (dx/2) source_vol, see the final scaling)
  {
    double f1[4][M][M], f2[4][M][M];
#pragma unroll
    for (int qx = 0; qx < M; ++qx)
#pragma unroll
      for (int qy = 0; qy < M; ++qy) {
        const fastm::Prim w = fastm::prim(P, U[0][qx][qy], U[1][qx][qy], U[2][qx][qy], U[3][qx][qy]);
        const double t = w.w0 * w.vx * w.vy, Ep = U[3][qx][qy] + w.p;
        f1[0][qx][qy] = w.w0 * w.vx; f1[1][qx][qy] = fma(w.vx, U[1][qx][qy], w.p); f1[2][qx][qy] = t; f1[3][qx][qy] = w.vx * Ep;
        f2[0][qx][qy] = w.w0 * w.vy; f2[1][qx][qy] = t; f2[2][qx][qy] = fma(w.vy, U[2][qx][qy], w.p); f2[3][qx][qy] = w.vy * Ep;
        // source at the node (get_source :1558-1576 / get_adv_source :1579-1596), kept in U (no longer needed)
        if (P.source == 2) {
          const double g1 = gx[(size_t)(qy * M + qx) * g.ne + e], g2 = gy[(size_t)(qy * M + qx) * g.ne + e];
          U[0][qx][qy] = 0.0; U[1][qx][qy] = w.w0 * g1; U[2][qx][qy] = w.w0 * g2; U[3][qx][qy] = w.w0 * fma(w.vx, g1, w.vy * g2);
        } else if (P.source == 3) {
          U[0][qx][qy] = -U[0][qx][qy]; U[1][qx][qy] = 0.0; U[2][qx][qy] = 0.0; U[3][qx][qy] = 0.0;
        }
      }
    const double src_scale = 0.5 * P.dx;     // source_vol/4 relative to the oneoverdx/2 scaling applied to acc below
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      // vol1(a,b) = sum_qx sum_qy f1 dPw[qx][a] Pw[qy][b];  vol2(a,b) = sum f2 Pw[qx][a] dPw[qy][b]
      double g1[M][M], g2[M][M];      // [a][qy]
#pragma unroll
      for (int a = 0; a < M; ++a)
#pragma unroll
        for (int qy = 0; qy < M; ++qy) {
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int qx = 0; qx < M; ++qx) {
            if (a > 0 && !zD<M>(qx, a)) s1 = fma(f1[v][qx][qy], B.dPw[qx][a], s1);      // dPw[.][0] = 0
            if (!zP<M>(qx, a)) s2 = fma(f2[v][qx][qy], B.Pw[qx][a], s2);
          }
          g1[a][qy] = s1; g2[a][qy] = s2;
        }
#pragma unroll
      for (int a = 0; a < M; ++a)
#pragma unroll
        for (int b = 0; b < M; ++b) {
          double s = acc[v][a][b];
#pragma unroll
          for (int qy = 0; qy < M; ++qy) {
            if (b > 0 && !zD<M>(qy, b)) s = fma(g2[a][qy], B.dPw[qy][b], s);
            if (a > 0 && !zP<M>(qy, b)) s = fma(g1[a][qy], B.Pw[qy][b], s);
          }
          acc[v][a][b] = s;          // vol1 + vol2 - (e1-e2) - (e3-e4)
        }
      if (P.source != 1) {
        // source_vol(a,b) = sum S Pw[qx][a] Pw[qy][b]; enters dudt as /4 while the flux terms enter as oneoverdx/2
#pragma unroll
        for (int a = 0; a < M; ++a)
#pragma unroll
          for (int qy = 0; qy < M; ++qy) {
            double s = 0.0;
#pragma unroll
            for (int qx = 0; qx < M; ++qx)
              if (!zP<M>(qx, a)) s = fma(U[v][qx][qy], B.Pw[qx][a], s);
            g1[a][qy] = s;
          }
#pragma unroll
        for (int a = 0; a < M; ++a)
#pragma unroll
          for (int b = 0; b < M; ++b) {
            double s = 0.0;
#pragma unroll
            for (int qy = 0; qy < M; ++qy)
              if (!zP<M>(qy, b)) s = fma(g1[a][qy], B.Pw[qy][b], s);
            acc[v][a][b] = fma(src_scale, s, acc[v][a][b]);
          }
      }
    }
  }
  // ---- dudt = (odx*vol1 + odx*vol2 - odx*(e1-e2) - odx*(e3-e4))/2 + source_vol/4      (:1449-1466)
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int a = 0; a < M; ++a)
#pragma unroll
      for (int b = 0; b < M; ++b) {
        double r = (0.5 * P.oneoverdx) * acc[v][a][b];
        if (fz && fz[(size_t)(b * M + a) * g.ne + e]) r = 0.0;
        acc[v][a][b] = r;
      }

  // ---- RK combination (real(4) coefficients of :683-707), in registers
  const double cdt = C.cd * ctrl->dt;
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int b = 0; b < M; ++b)
#pragma unroll
      for (int a = 0; a < M; ++a) {
        const int m = b * M + a;
        const double L = acc[v][a][b];
        if (C.out2) {     // everything of out2 that does not depend on the limited `out` (U is dead by now)
          double r2 = C.k0 * PL(C.B0, g, v, m)[e];
          r2 = fma(C.k1, PL(C.B1, g, v, m)[e], r2);
          r2 = fma(C.k2, src.rk_in(v, m), r2);
          U[v][a][b] = fma(C.ke * ctrl->dt, L, r2);
        }
        const double a0 = src.rk_a0(C, v, m);
        double r = (C.c0 == 1.0) ? a0 : C.c0 * a0;
        if (C.na >= 2) r = fma(C.c1, src.rk_a1(C, v, m), r);
        acc[v][a][b] = fma(cdt, L, r);
      }
  if (apply_onp) positivity_fast<M>(P, B, acc);
  if (C.out2) {
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int b = 0; b < M; ++b)
#pragma unroll
        for (int a = 0; a < M; ++a) PL(C.out2, g, v, b * M + a)[e] = fma(C.k3, acc[v][a][b], U[v][a][b]);
  }
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int b = 0; b < M; ++b)
#pragma unroll
      for (int a = 0; a < M; ++a) PL(out, g, v, b * M + a)[e] = acc[v][a][b];
}

// A step enqueued past `tend` is skipped on the device, but the host has already rotated its buffer pointers for it
// (new delta_u = the last stage's `out`): a skipped stage therefore hands its input through unchanged, so that after
// any number of skipped steps the buffer the host calls delta_u holds delta_u.
template <int M>
__device__ __forceinline__ void dg_stage_pass_through(const double* __restrict__ in, double* __restrict__ out, const DgGrid& g, size_t e) {
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int m = 0; m < M * M; ++m) PL(out, g, v, m)[e] = PL(in, g, v, m)[e];
}

template <int M, bool ANYFLUX>
__global__ void __launch_bounds__(64) k_dg_stage_fast(const double* __restrict__ in, StageCoef C, double* __restrict__ out,
                                                      const double* __restrict__ gx, const double* __restrict__ gy,
                                                      const unsigned char* __restrict__ fz, DgGrid g, DgPhys P, FastBasis B,
                                                      const DgCtrl* __restrict__ ctrl, int apply_onp) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne) return;
  if (ctrl->skip) { dg_stage_pass_through<M>(in, out, g, e); return; }
  const int ic = (int)(e % g.nx), jc = (int)(e / g.nx);
  GlobalSrc<M> src{in, g, e,
                   {(size_t)jc * g.nx + bc_index(P.bc, ic - 1, g.nyg), (size_t)jc * g.nx + bc_index(P.bc, ic + 1, g.nyg),
                    (size_t)y_nb(g, P.bc, jc - 1) * g.nx + ic, (size_t)y_nb(g, P.bc, jc + 1) * g.nx + ic}};
  dg_stage_body<M, ANYFLUX>(src, in, C, out, gx, gy, fz, g, P, B, ctrl, apply_onp, e);
}

}}  // namespace wb::dg
