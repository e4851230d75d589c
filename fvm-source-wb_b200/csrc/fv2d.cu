// 2D well-balanced finite-volume RK-stage kernels + C-ABI (replaces benchmark_2d.f90:221-279,370-618).
//
// Device layout: structure-of-arrays planes  plane_v[(r)*pitch + i],  v = rho, mx, my, E,
// i contiguous (x), r = local row + 1 (row 0 and row nyl+1 are the slab ghost rows).
//
// Kernels
//   k_stage_ref<MODE,WB>   reference operation order (arith = 1, parity triage, plain compute_update)
//   k_stage_march<MODE>    fused RK-stage kernel (production): warps march up strips of rows, each face flux
//                          computed once (warp shuffles in x, register carry in y), well-balanced source, RK
//                          axpy and (stage 2) the warp-shuffle max reduction for the next CFL time step.
//   (Two earlier designs were measured and dropped: a shared-memory tiled kernel -- 2x slower, barrier and
//    load-latency bound at 16 warps/SM -- and a two-cells-per-thread march -- same speed at 252 registers.)
//   MODE 0: out = dudt      MODE 1: out = in + dt*dudt      MODE 2: out = .5*base + .5*in + .5*dt*dudt
#include "common.cuh"
#include "fv2d_math.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace wb { namespace fv2d {

struct Ctrl {
  unsigned long long cmax_bits[2];  // max wave speed (bit pattern of a non-negative double), per step parity
  double t[2];
  int iter[2];
  double dt_last, cmax_last;
  int nonzero_vel;                  // set when the supplied w_eq has a non-zero velocity somewhere
};

struct Grid {
  int nx, ny;        // global
  int j0, nyl;       // first global row of this slab, local row count
  int pitch;         // doubles per row
  size_t plane;      // doubles per plane ((nyl+2)*pitch)
};

struct StageArgs {
  const double* in;     // state whose RHS is evaluated (4 planes)
  const double* base;   // MODE 2: u^n (4 planes)
  double* out;          // 4 planes
  const double* weq;    // ref kernels: supplied primitive equilibrium at centres (4 planes)
  const double* eqz;    // (rho_e, E_e) planes at the cell centres: conversions to / from the delta form, exact CFL speed
  int eq_exact;         // fused kernels: 1 = the supplied centre equilibrium is not the analytic one, read eqz for the CFL speed
  const double* exf; const double* exc; const double* eyf; const double* eyc;  // separable exp tables
  Ctrl* ctrl;
  int parity;
  double tend;
  int max_iter;
  int row_begin, row_end;   // local rows [row_begin,row_end) covered by this launch
  int rows_cap;             // TMA kernel: rows a strip computes (normally its stride R; 1 for the slab's two boundary rows, which
                            // travel as ONE launch with stride nyl-1: strip 0 = row 0, strip 1 = row nyl-1)
  int pf_rows;              // marching kernel: L2 prefetch distance in rows (0 = off)
  // slab boundary-row launch with peer-memory ghost rows: the row's results are ALSO stored into the neighbour's ghost row
  // (row 0 -> top ghost row of the slab below, row nyl-1 -> bottom ghost row of the slab above), over NVLink
  double* peer_lo; double* peer_hi;        // element (ghost row, column 0) of plane 0 of the neighbour's output field, or null
  size_t peer_lo_plane, peer_hi_plane;     // the neighbour's plane stride (its slab may be one row taller)
};

// ------------------------------------------------------------------------------------ layout kernels
// eqz != nullptr: the planes hold the DELTA FORM (u - u_eq at the cell centre, benchmark_2d.f90:499; the momenta of u_eq are
// zero), which is what the fused stage kernels keep resident -- the subtraction happens once here instead of in every stage
__global__ void k_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, Grid g, const double* __restrict__ eqz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y;
  if (i >= g.nx) return;
  const double2* src = reinterpret_cast<const double2*>(aos + ((size_t)j * g.nx + i) * 4);
  double2 a = src[0], b = src[1];
  size_t o = (size_t)(j + 1) * g.pitch + i;
  if (eqz) { a.x = a.x - eqz[o]; b.y = b.y - eqz[g.plane + o]; }
  soa[o] = a.x;
  soa[g.plane + o] = a.y;
  soa[2 * g.plane + o] = b.x;
  soa[3 * g.plane + o] = b.y;
}
// in place, all rows incl. ghosts: u -> u - u_eq (device-initialised states)
__global__ void k_to_delta(double* __restrict__ u, const double* __restrict__ eqz, Grid g) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (i >= g.nx) return;
  size_t o = (size_t)r * g.pitch + i;
  u[o] = u[o] - eqz[o];
  u[3 * g.plane + o] = u[3 * g.plane + o] - eqz[g.plane + o];
}
__global__ void k_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, Grid g, const double* __restrict__ eqz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y;
  if (i >= g.nx) return;
  size_t o = (size_t)(j + 1) * g.pitch + i;
  double2 a = make_double2(soa[o], soa[g.plane + o]);
  double2 b = make_double2(soa[2 * g.plane + o], soa[3 * g.plane + o]);
  if (eqz) { a.x = eqz[o] + a.x; b.y = eqz[g.plane + o] + b.y; }      // u = u_eq + delta
  double2* dst = reinterpret_cast<double2*>(aos + ((size_t)j * g.nx + i) * 4);
  dst[0] = a;
  dst[1] = b;
}

// (rho_e, E_e) planes from the supplied primitive equilibrium, all rows incl. ghosts
// (compute_conservative, benchmark_2d.f90:159-171, :496); flags non-zero equilibrium velocity.
__global__ void k_prepare_eq(const double* __restrict__ weq, double* __restrict__ eqz, Grid g, Phys P,
                             Ctrl* ctrl, const double* __restrict__ exc, const double* __restrict__ eyc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (i >= g.nx) return;
  size_t o = (size_t)r * g.pitch + i;
  double w[4] = {weq[o], weq[g.plane + o], weq[2 * g.plane + o], weq[3 * g.plane + o]};
  double u[4];
  ref::cons(P, w, u);
  eqz[o] = u[0];
  eqz[g.plane + o] = u[3];
  int jg = g.j0 + r - 1;
  if ((w[1] != 0.0 || w[2] != 0.0) && jg >= 0 && jg < g.ny) atomicOr(&ctrl->nonzero_vel, 1);
  // is the supplied centre equilibrium the analytic one (to rounding)?  The fused stage-2 kernel then rebuilds it from the
  // separable tables for the CFL speed; otherwise it reads these planes (bit 1 of the flag)
  if (r >= 1 && r <= g.nyl) {
    const double e = exc[i] * eyc[r - 1];
    if (!(fabs(u[0] - P.rho0 * e) <= 1e-9 * fabs(u[0])) || !(fabs(u[3] - P.pe1 * e) <= 1e-9 * fabs(u[3]))) atomicOr(&ctrl->nonzero_vel, 2);
  }
}

// ------------------------------------------------------------------------------------ IC on device
// get_initial_conditions (benchmark_2d.f90:45-113) and the centre equilibrium (:174-218) for rows
// r = 0..nyl+1 that exist globally.
__global__ void k_init(double* __restrict__ u, double* __restrict__ weq, Grid g, Phys P, int ninit,
                       double eta, int fill_u, int fill_weq) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (i >= g.nx) return;
  int jg = g.j0 + r - 1;
  size_t o = (size_t)r * g.pitch + i;
  if (jg < 0 || jg >= g.ny) {
    if (fill_u) { u[o] = 1.0; u[g.plane + o] = 0.0; u[2 * g.plane + o] = 0.0; u[3 * g.plane + o] = 1.0; }
    if (fill_weq) { weq[o] = 1.0; weq[g.plane + o] = 0.0; weq[2 * g.plane + o] = 0.0; weq[3 * g.plane + o] = 1.0; }
    return;
  }
  double x = x_cent(i, P.dx), y = x_cent(jg, P.dy);
  if (fill_weq) {
    double rho, p;
    ref::eq_prim(P, x, y, rho, p);
    weq[o] = rho; weq[g.plane + o] = 0.0; weq[2 * g.plane + o] = 0.0; weq[3 * g.plane + o] = p;
  }
  if (fill_u) {
    double w[4];
    const double rho_0 = (double)1.21f;
    if (ninit == 1) {
      double e = exp(-(x + y));
      w[0] = e; w[1] = 0; w[2] = 0; w[3] = e;
    } else if (ninit == 2 || ninit == 3) {
      double e = exp(-(rho_0 * 1.0 / 1.0) * (x + y));
      w[0] = rho_0 * e; w[1] = 0; w[2] = 0; w[3] = 1.0 * e;
      if (ninit == 3) {
        double ddx = x - (double)0.3f, ddy = y - (double)0.3f;
        w[3] = w[3] + eta * exp(-(100.0 * (rho_0 * 1.0 / 1.0) * (ddx * ddx + ddy * ddy)));
      }
    } else {
      if (x >= 0.5 && y >= 0.5)      { w[0] = 1.5; w[1] = 0.; w[2] = 0.; w[3] = 1.5; }
      else if (x < 0.5 && y >= 0.5)  { w[0] = (double)0.5323f; w[1] = (double)1.206f; w[2] = 0.; w[3] = (double)0.3f; }
      else if (x < 0.5 && y < 0.5)   { w[0] = (double)0.138f; w[1] = (double)1.206f; w[2] = (double)1.206f; w[3] = (double)0.029f; }
      else                           { w[0] = (double)0.5323f; w[1] = 0.; w[2] = (double)1.206f; w[3] = (double)0.3f; }
    }
    double c[4];
    ref::cons(P, w, c);
    u[o] = c[0]; u[g.plane + o] = c[1]; u[2 * g.plane + o] = c[2]; u[3 * g.plane + o] = c[3];
  }
}

// ------------------------------------------------------------------------------------ max speed
// compute_max_speed (benchmark_2d.f90:264-279): max over all cells (boundary included) of
// sqrt(vx^2+vy^2) + sqrt(gamma*max(p,1d-10)/max(rho,1d-10)); warp-shuffle + one atomicMax per block.
template <bool FAST>
__global__ void k_max_speed(const double* __restrict__ u, Grid g, Phys P, unsigned long long* out, const double* __restrict__ eqz) {
  double m = 0.0;
  for (int r = blockIdx.y + 1; r <= g.nyl; r += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.nx; i += gridDim.x * blockDim.x) {
      size_t o = (size_t)r * g.pitch + i;
      double s;
      if (FAST) {
        double u0 = u[o], u3 = u[3 * g.plane + o];
        if (eqz) { u0 = eqz[o] + u0; u3 = eqz[g.plane + o] + u3; }      // delta form
        s = fast::speed(P, u0, u[g.plane + o], u[2 * g.plane + o], u3);
      } else {
        double uu[4] = {u[o], u[g.plane + o], u[2 * g.plane + o], u[3 * g.plane + o]};
        s = ref::speed(P, uu);
      }
      m = fmax(m, s);
    }
  m = warp_max(m);
  __shared__ double sm[32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sm[w] = m;
  __syncthreads();
  if (w == 0) {
    m = (lane < (blockDim.x >> 5)) ? sm[lane] : 0.0;
    m = warp_max(m);
    if (lane == 0) atomic_max_nonneg(out, m);
  }
}

// output_file (benchmark_2d.f90:115-143): one row (x, y, p - p_eq) per cell, icell outer / jcell inner, from the resident
// state (delta form when eqz != nullptr); equilibrium by get_equilibrium_solution at the centre, pressure by compute_primitive
__global__ void k_pack_output(const double* __restrict__ u, const double* __restrict__ eqz, Grid g, Phys P, double* __restrict__ tab) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y;
  if (i >= g.nx) return;
  const size_t o = (size_t)(j + 1) * g.pitch + i;
  double uu[4] = {u[o], u[g.plane + o], u[2 * g.plane + o], u[3 * g.plane + o]}, w[4];
  if (eqz) { uu[0] = eqz[o] + uu[0]; uu[3] = eqz[g.plane + o] + uu[3]; }
  ref::prim(P, uu, w);
  const double x = x_cent(i, P.dx), y = x_cent(g.j0 + j, P.dy);
  double rho_e, p_e;
  ref::eq_prim(P, x, y, rho_e, p_e);
  double* row = tab + ((size_t)i * g.nyl + j) * 3;
  row[0] = x; row[1] = y; row[2] = w[3] - p_e;
}

__global__ void k_ctrl_reset(Ctrl* c) {
  c->cmax_bits[0] = 0ull; c->cmax_bits[1] = 0ull;
  c->t[0] = 0.0; c->t[1] = 0.0;
  c->iter[0] = 0; c->iter[1] = 0;
  c->dt_last = 0.0; c->cmax_last = 0.0;
}

__device__ __forceinline__ bool step_done(const Ctrl* c, int parity, double tend, int max_iter) {
  double t = c->t[parity];
  int it = c->iter[parity];
  return !(t < tend) || (max_iter >= 0 && it >= max_iter);
}
// dt = 0.5*dx/cmax*cfl   (benchmark_2d.f90:242)
__device__ __forceinline__ double step_dt(const Ctrl* c, int parity, const Phys& P) {
  double cmax = __longlong_as_double((long long)c->cmax_bits[parity]);
  return 0.5 * P.dx / cmax * P.cfl;
}
// a finished run (t >= tend) turns the remaining enqueued steps into no-ops; the stage-2 launch
// carries the bookkeeping slot forward so that the next parity sees the same (t, iter, cmax)
__device__ __forceinline__ void carry_forward(Ctrl* c, int parity) {
  c->t[parity ^ 1] = c->t[parity];
  c->iter[parity ^ 1] = c->iter[parity];
  c->cmax_bits[parity ^ 1] = c->cmax_bits[parity];
}
template <int MODE>
__device__ __forceinline__ void bookkeeping(Ctrl* c, int parity, double dt) {
  if (MODE == 1) c->cmax_bits[parity ^ 1] = 0ull;
  if (MODE == 2) {
    c->t[parity ^ 1] = c->t[parity] + dt;
    c->iter[parity ^ 1] = c->iter[parity] + 1;
    c->dt_last = dt;
    c->cmax_last = __longlong_as_double((long long)c->cmax_bits[parity]);
  }
}

// ------------------------------------------------------------------------------------ reference-order stage
// One thread per cell, every quantity recomputed from global memory exactly as
// compute_update_exact does (benchmark_2d.f90:465-618); WB=false is the plain compute_update (:370-463).
template <int MODE, bool WB>
__global__ void __launch_bounds__(128) k_stage_ref(StageArgs A, Grid g, Phys P) {
  double dt = 0.0;
  if (MODE != 0) {
    if (step_done(A.ctrl, A.parity, A.tend, A.max_iter)) {
      if (MODE == 2 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && A.row_begin == 0)
        carry_forward(A.ctrl, A.parity);
      return;
    }
    dt = step_dt(A.ctrl, A.parity, P);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && A.row_begin == 0)
      bookkeeping<MODE>(A.ctrl, A.parity, dt);
  }
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int jl = A.row_begin + blockIdx.y;
  double spd = 0.0;
  if (i < g.nx && jl < A.row_end) {
    int jg = g.j0 + jl;
    size_t o = (size_t)(jl + 1) * g.pitch + i;
    double uc[4] = {A.in[o], A.in[g.plane + o], A.in[2 * g.plane + o], A.in[3 * g.plane + o]};
    double d[4] = {0.0, 0.0, 0.0, 0.0};
    bool interior = (i > 0 && i < g.nx - 1 && jg > 0 && jg < g.ny - 1);
    if (interior) {
      // neighbours: left, right, bottom, top
      const size_t on[4] = {o - 1, o + 1, o - g.pitch, o + g.pitch};
      double dc[4], dn[4][4];
      if (WB) {
        double wq[4] = {A.weq[o], A.weq[g.plane + o], A.weq[2 * g.plane + o], A.weq[3 * g.plane + o]};
        double ue[4];
        ref::cons(P, wq, ue);
#pragma unroll
        for (int v = 0; v < 4; ++v) dc[v] = uc[v] - ue[v];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          double wn[4] = {A.weq[on[k]], A.weq[g.plane + on[k]], A.weq[2 * g.plane + on[k]], A.weq[3 * g.plane + on[k]]};
          double un[4];
          ref::cons(P, wn, un);
#pragma unroll
          for (int v = 0; v < 4; ++v) dn[k][v] = A.in[v * g.plane + on[k]] - un[v];
        }
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v) dc[v] = uc[v];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int v = 0; v < 4; ++v) dn[k][v] = A.in[v * g.plane + on[k]];
      }
      // conservative equilibrium at the four faces of the cell
      double UX[2][4], UY[2][4];
      if (WB) {
        double xc = x_cent(i, P.dx), yc = x_cent(jg, P.dy);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          double w[4] = {0, 0, 0, 0};
          ref::eq_prim(P, x_face(i + k, P.dx), yc, w[0], w[3]);
          ref::cons(P, w, UX[k]);
          ref::eq_prim(P, xc, x_face(jg + k, P.dx) /* (j-1)*dx sic, :513 */, w[0], w[3]);
          ref::cons(P, w, UY[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int v = 0; v < 4; ++v) { UX[k][v] = 0.0; UY[k][v] = 0.0; }
      }
      double ul[4], ur[4], F0[4], F1[4], G0[4], G1[4];
      // x-face i: left state u_right(i-1), right state u_left(i)
#pragma unroll
      for (int v = 0; v < 4; ++v) { ul[v] = WB ? UX[0][v] + dn[0][v] : dn[0][v]; ur[v] = WB ? UX[0][v] + dc[v] : dc[v]; }
      ref::llf<0>(P, ul, ur, F0);
#pragma unroll
      for (int v = 0; v < 4; ++v) { ul[v] = WB ? UX[1][v] + dc[v] : dc[v]; ur[v] = WB ? UX[1][v] + dn[1][v] : dn[1][v]; }
      ref::llf<0>(P, ul, ur, F1);
      // y-face j: lower state u_top(j-1), upper state u_bottom(j)
#pragma unroll
      for (int v = 0; v < 4; ++v) { ul[v] = WB ? UY[0][v] + dn[2][v] : dn[2][v]; ur[v] = WB ? UY[0][v] + dc[v] : dc[v]; }
      ref::llf<1>(P, ul, ur, G0);
#pragma unroll
      for (int v = 0; v < 4; ++v) { ul[v] = WB ? UY[1][v] + dc[v] : dc[v]; ur[v] = WB ? UY[1][v] + dn[3][v] : dn[3][v]; }
      ref::llf<1>(P, ul, ur, G1);
      double w[4], s[4];
      ref::prim(P, uc, w);
      ref::source(w, s);
      if (WB) {
        double wq[4] = {A.weq[o], A.weq[g.plane + o], A.weq[2 * g.plane + o], A.weq[3 * g.plane + o]};
        double se[4], Fe0[4], Fe1[4], Ge0[4], Ge1[4];
        ref::source(wq, se);
        ref::flux<0>(P, UX[0], Fe0);
        ref::flux<0>(P, UX[1], Fe1);
        ref::flux<1>(P, UY[0], Ge0);
        ref::flux<1>(P, UY[1], Ge1);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          double r = -(F1[v] - F0[v]) * P.odx - (G1[v] - G0[v]) * P.ody;
          r = r + s[v];
          r = r - se[v];
          r = r + (Fe1[v] - Fe0[v]) * P.odx;
          r = r + (Ge1[v] - Ge0[v]) * P.ody;
          d[v] = r;
        }
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v) d[v] = -(F1[v] - F0[v]) * P.odx - (G1[v] - G0[v]) * P.ody + s[v];
      }
    }
    double un[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (MODE == 0) un[v] = d[v];
      if (MODE == 1) un[v] = uc[v] + dt * d[v];                                   // w1=u+dt*dudt  :247
      if (MODE == 2) un[v] = 0.5 * A.base[v * g.plane + o] + 0.5 * uc[v] + 0.5 * dt * d[v];  // :250
      A.out[v * g.plane + o] = un[v];
    }
    if (MODE == 2) spd = ref::speed(P, un);
  }
  if (MODE == 2) {
    spd = warp_max(spd);
    __shared__ double sm[4];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = spd;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = fmax(fmax(sm[0], sm[1]), fmax(sm[2], sm[3]));
      atomic_max_nonneg(&A.ctrl->cmax_bits[A.parity ^ 1], m);
    }
  }
}

// ------------------------------------------------------------------------------------ fused stage: face flux helpers
// One LLF face in (normal,tangential) form.  lo/hi = delta on the low/high side of the face,
// (rf,Ef) = conservative equilibrium at the face.  Returns TWICE the flux (mass, normal mom, tangential mom, energy)
// MINUS twice the equilibrium flux of the face, whose only non-zero entry is the pressure 2*gm1*Ef in the normal
// momentum: at delta = 0 both sides carry p = gm1*Ef from the same instruction sequence, so fn is exactly zero there
// and the "+ (F_eq(i+1)-F_eq(i))/dx" terms of benchmark_2d.f90:605-606 need no separate evaluation.
struct FaceFlux { double f0, fn, ft, f3; };
__device__ __forceinline__ FaceFlux face_llf(const Phys& P, double rf, double Ef, double lr, double ln,
                                             double lt, double lE, double hr, double hn, double ht,
                                             double hE) {
  // states: equilibrium at the face + the cell's own delta (benchmark_2d.f90:533-537)
  double ar = rf + lr, aE = Ef + lE;   // low side  ("u_right(i-1)" / "u_top(j-1)")
  double br = rf + hr, bE = Ef + hE;   // high side ("u_left(i)"   / "u_bottom(j)")
  fast::Eval a = fast::eval_state(P, ar, ln, lt, aE);
  fast::Eval b = fast::eval_state(P, br, hn, ht, bE);
  const double cm = fast::max_speed2(a.spd, b.spd);
  FaceFlux o;
  // 2 x [0.5*(f_right+f_left)+0.5*cmax*(uleft-uright)]   (benchmark_2d.f90:366): the two halvings are exact, so they
  // are folded into 0.5/dx (Phys::hodx, hody); the remaining a*b+c is one fma (<= 1 ulp of an O(1) flux)
  o.f0 = fma(cm, ar - br, b.f0 + a.f0);
  o.fn = fma(cm, ln - hn, (b.fn + a.fn) - P.gm1x2 * Ef);
  o.ft = fma(cm, lt - ht, b.ft + a.ft);
  o.f3 = fma(cm, aE - bE, b.f3 + a.f3);
  return o;
}

// Two LLF faces at once: the four states are evaluated in lock-step (eval_states<4>); same arithmetic as face_llf.
struct FaceIn { double rf, Ef, lr, ln, lt, lE, hr, hn, ht, hE; };
template <bool EXACT>
__device__ __forceinline__ void faces_llf2(const Phys& P, const FaceIn& a, const FaceIn& b, FaceFlux& oa, FaceFlux& ob, bool& ok) {
  const double rho[4] = {a.rf + a.lr, a.rf + a.hr, b.rf + b.lr, b.rf + b.hr};
  const double E[4] = {a.Ef + a.lE, a.Ef + a.hE, b.Ef + b.lE, b.Ef + b.hE};
  const double mn[4] = {a.ln, a.hn, b.ln, b.hn};
  const double mt[4] = {a.lt, a.ht, b.lt, b.ht};
  fast::Eval ev[4];
  fast::eval_states<4, EXACT>(P, rho, mn, mt, E, ev, ok);
  const double cma = fast::max_speed2(ev[0].spd, ev[1].spd), cmb = fast::max_speed2(ev[2].spd, ev[3].spd);
  oa.f0 = fma(cma, rho[0] - rho[1], ev[1].f0 + ev[0].f0);
  ob.f0 = fma(cmb, rho[2] - rho[3], ev[3].f0 + ev[2].f0);
  oa.fn = fma(cma, a.ln - a.hn, (ev[1].fn + ev[0].fn) - P.gm1x2 * a.Ef);
  ob.fn = fma(cmb, b.ln - b.hn, (ev[3].fn + ev[2].fn) - P.gm1x2 * b.Ef);
  oa.ft = fma(cma, a.lt - a.ht, ev[1].ft + ev[0].ft);
  ob.ft = fma(cmb, b.lt - b.ht, ev[3].ft + ev[2].ft);
  oa.f3 = fma(cma, E[0] - E[1], ev[1].f3 + ev[0].f3);
  ob.f3 = fma(cmb, E[2] - E[3], ev[3].f3 + ev[2].f3);
}

// ------------------------------------------------------------------------------------ fused marching stage
// The production RK-stage kernel.  No shared memory, no block barriers:
//   * a warp owns 32 consecutive columns and marches up a strip of rows; lane l evaluates the LEFT x-face
//     and the TOP y-face of its cell in every row, so each face flux is computed exactly once;
//   * the right x-face comes from lane l+1 by warp shuffle (lane 31 only feeds lane 30: 31 outputs / warp),
//     the left neighbour's delta from lane l-1 (lane 0 loads its halo column itself);
//   * the bottom y-face flux is carried in registers from the previous row;
//   * rows j+2 are prefetched while row j is computed (software pipelining instead of occupancy).
// HBM traffic per cell: read 4 doubles [+ u^n (4) in stage 2], write 4 (delta form, see Cell).
constexpr int MARCH_WARPS = 4;      // warps per CTA (independent of each other)
constexpr int MARCH_OUT = 31;       // output columns per warp
#ifndef MARCH_MIN_BLOCKS
#define MARCH_MIN_BLOCKS 4
#endif

// The fused kernels keep the state in DELTA FORM: planes (rho - rho_e, mx, my, E - E_e) with (rho_e, E_e) the conservative
// equilibrium at the cell centre (benchmark_2d.f90:496-499; its momenta are zero).  The subtraction is done once at upload
// (k_aos_to_soa / k_to_delta) and undone at download, so a stage reads 32 B and writes 32 B per cell and nothing else:
// the face equilibria come from the separable exp tables, the source term  s - s_eq = -(rho - rho_e)  is the density
// perturbation itself, and the RK combination is linear, so it acts on the perturbation unchanged.
struct Cell { double d0, d1, d2, d3; };

__device__ __forceinline__ Cell load_cell(const StageArgs& A, const Grid& g, size_t o) {
  Cell c;
  c.d0 = A.in[o]; c.d1 = A.in[g.plane + o]; c.d2 = A.in[2 * g.plane + o]; c.d3 = A.in[3 * g.plane + o];
  return c;
}

// Max wave speed of the updated cell (stage-2 CFL reduction, benchmark_2d.f90:264-295) from its delta form: the centre
// equilibrium is rho0*e, p0/(gamma-1)*e with e = exp(-a xc) exp(-a yc) from the separable tables -- or, when the caller's
// w_eq is not the analytic equilibrium, the planes it was given.
// (out of line: the rare case must not cost the common one ten predicated-off instructions per cell)
__device__ __noinline__ double2 centre_eq_supplied(const double* eqz, size_t plane, size_t o) {
  return make_double2(eqz[o], eqz[plane + o]);
}
__device__ __forceinline__ double centre_speed(const StageArgs& A, const Grid& g, const Phys& P, size_t o, double e, double n0,
                                               double n1, double n2, double n3) {
  double re = P.rho0 * e, Ee = P.pe1 * e;
  if (A.eq_exact) { const double2 q = centre_eq_supplied(A.eqz, g.plane, o); re = q.x; Ee = q.y; }
  return fast::speed(P, re + n0, n1, n2, Ee + n3);
}

// dudt of one cell from its four face fluxes, in the reference's order (benchmark_2d.f90:601-607)
template <int MODE>
__device__ __forceinline__ void cell_update(const Phys& P, const Cell& c, const FaceFlux& Fl, const FaceFlux& Fr, const FaceFlux& Gb,
                                            const FaceFlux& Gt, bool interior, double dt, double b0, double b1, double b2,
                                            double b3, double& n0, double& n1, double& n2, double& n3) {
  // the face fluxes arrive doubled and relative to the equilibrium flux: hodx = 0.5/dx, hody = 0.5/dy (exact
  // scalings); every term below vanishes identically at the hydrostatic state.
  double d0 = fma(Gb.f0 - Gt.f0, P.hody, -((Fr.f0 - Fl.f0) * P.hodx));
  double d1 = fma(Gb.ft - Gt.ft, P.hody, -((Fr.fn - Fl.fn) * P.hodx));
  double d2 = fma(Gb.fn - Gt.fn, P.hody, -((Fr.ft - Fl.ft) * P.hodx));
  double d3 = fma(Gb.f3 - Gt.f3, P.hody, -((Fr.f3 - Fl.f3) * P.hodx));
  d1 = d1 - c.d0;                      // s - s_eq = -(rho - rho_e)   (benchmark_2d.f90:327-350, :603-604)
  d2 = d2 - c.d0;
  d3 = d3 - (c.d1 + c.d2);
  if (MODE == 0) {
    if (!interior) { d0 = 0.0; d1 = 0.0; d2 = 0.0; d3 = 0.0; }
    n0 = d0; n1 = d1; n2 = d2; n3 = d3;
  }
  // frozen boundary lines (:611-614): dudt = 0 there, selected (not multiplied by a zero time step: on a zero-filled halo
  // column d may be NaN)
  if (MODE == 1) {
    n0 = interior ? fma(dt, d0, c.d0) : c.d0; n1 = interior ? fma(dt, d1, c.d1) : c.d1;
    n2 = interior ? fma(dt, d2, c.d2) : c.d2; n3 = interior ? fma(dt, d3, c.d3) : c.d3;
  }
  if (MODE == 2) {
    const double hdt = 0.5 * dt;
    const double a0 = 0.5 * (b0 + c.d0), a1 = 0.5 * (b1 + c.d1), a2 = 0.5 * (b2 + c.d2), a3 = 0.5 * (b3 + c.d3);
    n0 = interior ? fma(hdt, d0, a0) : a0; n1 = interior ? fma(hdt, d1, a1) : a1;
    n2 = interior ? fma(hdt, d2, a2) : a2; n3 = interior ? fma(hdt, d3, a3) : a3;
  }
}

// L2 prefetch of one row segment of one plane (bulk prefetch: one warp-level instruction brings `bytes` contiguous
// bytes into L2; nothing is written, no completion to wait for).  16-byte aligned address and size.
__device__ __forceinline__ void l2_prefetch_bulk(const char* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Per-thread constants of the marching loop.
struct MarchCtx {
  int lane, ic, ih, jmin, jmax;
  bool writer, col_interior;
  double exf_i, exc_i, dt;
};

// One row of the march.  `cur` is the cell of row j, `nraw` holds row j+1 as loaded one call ago and is
// refilled with row j+2; `hraw` likewise for lane 0's halo column; `Gb` is the bottom face flux (in) and
// `Gt` the top face flux (out).  The caller alternates (cur,Gb) <-> (nxt,Gt) so that no register is moved.
template <int MODE>
__device__ __forceinline__ void march_row(const StageArgs& A, const Grid& g, const Phys& P, const MarchCtx& c, int j,
                                          const Cell& cur, Cell& nxt, Cell& nraw, Cell& hraw, const FaceFlux& Gb,
                                          FaceFlux& Gt, double& spd) {
  auto row_off = [&](int jj) { return (size_t)(max(c.jmin, min(jj, c.jmax)) + 1) * g.pitch; };
  // ---- rows loaded during the previous call become usable now ...
  nxt = nraw;
  const Cell hal = hraw;
  // ---- ... and this row's loads are issued before any arithmetic: row j+2, lane 0's halo of row j+1, u^n of
  //      row j (stage 2).  They are consumed one row (~250 FP64 instructions) later.
  nraw = load_cell(A, g, row_off(j + 2) + c.ic);
  if (c.lane == 0) hraw = load_cell(A, g, row_off(j + 1) + c.ih);
  const size_t o = row_off(j) + c.ic;
  double b0 = 0, b1 = 0, b2 = 0, b3 = 0;
  if (MODE == 2) { b0 = A.base[o]; b1 = A.base[g.plane + o]; b2 = A.base[2 * g.plane + o]; b3 = A.base[3 * g.plane + o]; }

  // ---- left neighbour's delta (lane-1 / halo), then the top y-face (j+1; normal = y) and the left x-face (i;
  //      normal = x) together: their four states are evaluated in lock-step
  double l0 = __shfl_up_sync(0xffffffffu, cur.d0, 1), l1 = __shfl_up_sync(0xffffffffu, cur.d1, 1);
  double l2 = __shfl_up_sync(0xffffffffu, cur.d2, 1), l3 = __shfl_up_sync(0xffffffffu, cur.d3, 1);
  if (c.lane == 0) { l0 = hal.d0; l1 = hal.d1; l2 = hal.d2; l3 = hal.d3; }
  const double tyc = A.eyc[min(j, g.nyl - 1)];
  const double ey = c.exc_i * A.eyf[min(j + 1, g.nyl)];
  const double ex = c.exf_i * tyc;
  FaceFlux Fl;
  {
    const FaceIn fy = {P.rho0 * ey, P.pe1 * ey, cur.d0, cur.d2, cur.d1, cur.d3, nxt.d0, nxt.d2, nxt.d1, nxt.d3};
    const FaceIn fx = {P.rho0 * ex, P.pe1 * ex, l0, l1, l2, l3, cur.d0, cur.d1, cur.d2, cur.d3};
    bool ok = true;
    faces_llf2<true>(P, fy, fx, Gt, Fl, ok);
  }
  // ---- right x-face (i+1) from lane+1
  const double r0 = __shfl_down_sync(0xffffffffu, Fl.f0, 1), rn = __shfl_down_sync(0xffffffffu, Fl.fn, 1);
  const double rt = __shfl_down_sync(0xffffffffu, Fl.ft, 1), r3 = __shfl_down_sync(0xffffffffu, Fl.f3, 1);

  // ---- dudt in the reference's order (benchmark_2d.f90:601-607), RK axpy
  const int jg = g.j0 + j;
  const bool interior = c.col_interior && (jg > 0) && (jg < g.ny - 1);
  FaceFlux Fr;
  Fr.f0 = r0; Fr.fn = rn; Fr.ft = rt; Fr.f3 = r3;
  double n0, n1, n2, n3;
  cell_update<MODE>(P, cur, Fl, Fr, Gb, Gt, interior, c.dt, b0, b1, b2, b3, n0, n1, n2, n3);
  if (c.writer) {
    A.out[o] = n0; A.out[g.plane + o] = n1; A.out[2 * g.plane + o] = n2; A.out[3 * g.plane + o] = n3;
    if (MODE == 2) spd = fmax(spd, centre_speed(A, g, P, o, c.exc_i * tyc, n0, n1, n2, n3));
  }
}

template <int MODE, int MB>
__global__ void __launch_bounds__(MARCH_WARPS * 32, MB) k_stage_march(StageArgs A, Grid g, Phys P, int R) {
  MarchCtx c;
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
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // tells the compiler the value is warp-uniform
  const int wc = blockIdx.x * MARCH_WARPS + warp;
  const int c0 = wc * MARCH_OUT;
  if (c0 >= g.nx) return;                                  // whole warp: no barriers in this kernel
  const int jb = A.row_begin + blockIdx.y * R;
  const int je = min(jb + R, A.row_end);
  if (jb >= je) return;
  const int i = c0 + c.lane;
  c.ic = min(i, g.nx - 1);
  c.ih = max(c0 - 1, 0);                                   // lane 0's halo column
  // loadable local rows: ghost rows exist only where a neighbouring slab does
  c.jmin = (g.j0 > 0) ? -1 : 0;
  c.jmax = (g.j0 + g.nyl < g.ny) ? g.nyl : g.nyl - 1;
  auto row_off = [&](int j) { return (size_t)(max(c.jmin, min(j, c.jmax)) + 1) * g.pitch; };
  c.exf_i = A.exf[c.ic];
  c.exc_i = A.exc[c.ic];
  c.writer = (c.lane < MARCH_OUT) && (i < g.nx);
  c.col_interior = (i > 0) && (i < g.nx - 1);

  // ---- prologue: rows jb-1, jb, jb+1 and lane 0's halo of row jb; bottom face of the strip
  Cell ca = load_cell(A, g, row_off(jb) + c.ic), cb;
  Cell nraw = load_cell(A, g, row_off(jb + 1) + c.ic);
  Cell hraw = load_cell(A, g, row_off(jb) + (c.lane == 0 ? c.ih : c.ic));
  FaceFlux Ga, Gb2;
  {
    const Cell bel = load_cell(A, g, row_off(jb - 1) + c.ic);
    const double e = c.exc_i * A.eyf[jb];
    Ga = face_llf(P, P.rho0 * e, P.pe1 * e, bel.d0, bel.d2, bel.d1, bel.d3, ca.d0, ca.d2, ca.d1, ca.d3);
  }
  // ---- L2 prefetch stream.  The register prefetch above gives one row of lead, which at 4 warps per scheduler
  //      does not cover the DRAM latency (the first use of the prefetched row was 40 % of all stall samples).
  //      So every row each warp also asks L2 for the CTA's 1 KiB segment (4 x 31 columns + halo) of local row
  //      j + pf_rows of "its" input planes: plane p (u x4, u^n x4 in stage 2) belongs to warp p mod 4.
  //      UBLKPF is a uniform-datapath instruction: one issue per warp, addresses from warp-uniform registers.
  constexpr int NPF = (MODE == 2) ? 8 : 4;
  constexpr int NPW = (NPF + MARCH_WARPS - 1) / MARCH_WARPS;
  const char* pf[NPW];
  const size_t pitchB = (size_t)g.pitch * sizeof(double);
  unsigned pf_bytes = 0;
  int pf_last = -1;                                        // last local row worth prefetching for this strip
  if (A.pf_rows > 0) {
    const int s0 = max((int)blockIdx.x * MARCH_WARPS * MARCH_OUT - 1, 0) & ~1;   // even column: 16-byte aligned
    pf_bytes = (unsigned)(min(128, g.pitch - s0) * (int)sizeof(double));
    pf_last = min(je, c.jmax);
#pragma unroll
    for (int k = 0; k < NPW; ++k) {
      const int pl = min(warp + k * MARCH_WARPS, NPF - 1);
      const double* b = (pl < 4) ? A.in + (size_t)pl * g.plane : A.base + (size_t)(pl - 4) * g.plane;
      pf[k] = (const char*)(b + s0) + (size_t)(jb + A.pf_rows + 1) * pitchB;
    }
  }
  auto prefetch_row = [&](int jj) {                        // jj = local row to prefetch; all operands warp-uniform
    if (jj <= pf_last) {
#pragma unroll
      for (int k = 0; k < NPW; ++k)
        if (warp + k * MARCH_WARPS < NPF) l2_prefetch_bulk(pf[k], pf_bytes);
    }
#pragma unroll
    for (int k = 0; k < NPW; ++k) pf[k] += pitchB;
  };
  double spd = 0.0;
  int j = jb;
  for (; j + 1 < je; j += 2) {           // two rows per trip: (ca,Ga)->(cb,Gb2)->(ca,Ga), no register rotation
    prefetch_row(j + A.pf_rows);
    march_row<MODE>(A, g, P, c, j, ca, cb, nraw, hraw, Ga, Gb2, spd);
    prefetch_row(j + 1 + A.pf_rows);
    march_row<MODE>(A, g, P, c, j + 1, cb, ca, nraw, hraw, Gb2, Ga, spd);
  }
  if (j < je) march_row<MODE>(A, g, P, c, j, ca, cb, nraw, hraw, Ga, Gb2, spd);
  if (MODE == 2) {
    spd = warp_max(spd);
    if (c.lane == 0) atomic_max_nonneg(&A.ctrl->cmax_bits[A.parity ^ 1], spd);
  }
}


}}  // namespace wb::fv2d

#include "fv2d_tma.cuh"

// ============================================================================================ host side
using namespace wb;
using namespace wb::fv2d;

struct wb_fv2d {
  wb_fv2d_params prm;
  int dev = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  Grid g;
  Phys phys;
  double *u = nullptr, *w1 = nullptr, *weq = nullptr, *eqz = nullptr, *stage = nullptr, *tab = nullptr, *eqflag = nullptr;
  const double *exf = nullptr, *exc = nullptr, *eyf = nullptr, *eyc = nullptr;
  Ctrl* ctrl = nullptr;
  Ctrl* h_ctrl = nullptr;       // pinned
  bool resident = false;
  bool fast_ok = true;          // supplied equilibrium has zero velocity -> fused kernels usable
  bool eq_analytic = true;      // ... and is the analytic one at the cell centres (else the CFL speed reads the eqz planes)
  int parity = 0;
  wb::Nccl* comm = nullptr;
  cudaStream_t comm_stream = nullptr;   // slab mode: boundary rows + NCCL ghost exchange run here, overlapped with the interior
  cudaEvent_t ev_main = nullptr, ev_comm = nullptr;
  int overlap = 1;
  // peer-memory ghost rows (one process per GPU: the neighbours' state buffers and flag words are mapped with CUDA IPC)
  struct Peer {
    bool active = false;
    double *lo_u = nullptr, *lo_w1 = nullptr, *hi_u = nullptr, *hi_w1 = nullptr;      // neighbours' buffers (mapped)
    unsigned long long *lo_flags = nullptr, *hi_flags = nullptr;                      // neighbours' flag words (mapped)
    size_t lo_plane = 0, hi_plane = 0;
    int lo_nyl = 0, hi_nyl = 0;
    unsigned long long* flags = nullptr;      // own: [0] counts the rows received from below, [1] from above, [2] error word
    unsigned long long seq = 0;               // exchanges enqueued so far (the same number on every rank)
  } peer;
  int march_rows = 32;          // rows per strip of the marching kernel
  int pf_rows = 4;              // L2 prefetch distance of the LDG marching kernel (rows ahead; 0 = off)
  bool tma_ok = false;          // tensor maps built: the TMA-fed stage kernel is used
  CUtensorMap map_u, map_w1;
  wb::OutputJob* out_job = nullptr;   // output_file in flight (host thread)
};

namespace {

int fill_phys(const wb_fv2d_params& p, Phys& P) {
  P.gamma = p.gamma;
  P.gm1 = p.gamma - (double)1.0f;
  P.dx = p.boxlen_x / (double)p.nx;
  P.dy = p.boxlen_y / (double)p.ny;
  P.odx = 1 / P.dx;
  P.ody = 1 / P.dy;
  P.hodx = 0.5 * P.odx;
  P.hody = 0.5 * P.ody;
  P.cfl = p.cfl;
  P.neq = p.nequilibrium;
  if (p.nequilibrium == 1) { P.rho0 = 1.0; P.p0 = 1.0; P.a = 1.0; }
  else if (p.nequilibrium == 2 || p.nequilibrium == 3) {
    P.rho0 = (double)1.21f; P.p0 = 1.0; P.a = P.rho0 * 1.0 / 1.0;
  } else { P.rho0 = 0.0; P.p0 = 0.0; P.a = 0.0; }
  P.pe1 = P.p0 / P.gm1;
  P.gm1x2 = 2.0 * P.gm1;
  return WB_OK;
}

size_t stage_bytes(const wb_fv2d* h) { return sizeof(double) * 4 * (size_t)h->g.nx * h->g.nyl; }

int ensure_stage(wb_fv2d* h) {
  if (!h->stage) WB_CUDA(cudaMalloc(&h->stage, stage_bytes(h)));
  return WB_OK;
}

// host AoS (Fortran u(nvar,nx,ny_local)) -> device SoA planes
// (delta = true: stored as u - u_eq, the resident form of the fused kernels -- needs the eqz planes, i.e. prepare_eq first)
int h2d_state(wb_fv2d* h, const double* host, double* soa, bool delta = false) {
  WB_CHECK(ensure_stage(h));
  WB_CUDA(cudaMemcpyAsync(h->stage, host, stage_bytes(h), cudaMemcpyHostToDevice, h->stream));
  dim3 b(128), gr((h->g.nx + 127) / 128, h->g.nyl);
  k_aos_to_soa<<<gr, b, 0, h->stream>>>(h->stage, soa, h->g, delta ? h->eqz : nullptr);
  WB_LAUNCH_CHECK();
  return WB_OK;
}
int d2h_state(wb_fv2d* h, const double* soa, double* host, bool delta = false) {
  WB_CHECK(ensure_stage(h));
  dim3 b(128), gr((h->g.nx + 127) / 128, h->g.nyl);
  k_soa_to_aos<<<gr, b, 0, h->stream>>>(soa, h->stage, h->g, delta ? h->eqz : nullptr);
  WB_LAUNCH_CHECK();
  WB_CUDA(cudaMemcpyAsync(host, h->stage, stage_bytes(h), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  return WB_OK;
}

// ghost-row exchange of a 4-plane field with the slab neighbours (per-stage halo, NCCL send/recv)
int exchange_ghost_rows(wb_fv2d* h, double* field, int nplanes, cudaStream_t stream = nullptr) {
  if (!stream) stream = h->stream;
  if (h->prm.nranks <= 1) return WB_OK;
  if (!h->comm) { set_error("nranks > 1 but wb_fv2d_comm_init was not called"); return WB_ERR_STATE; }
  HaloSeg lo[4], hi[4];
  const Grid& g = h->g;
  for (int v = 0; v < nplanes; ++v) {
    double* p = field + v * g.plane;
    lo[v] = {p + (size_t)1 * g.pitch, p, (size_t)g.nx};
    hi[v] = {p + (size_t)g.nyl * g.pitch, p + (size_t)(g.nyl + 1) * g.pitch, (size_t)g.nx};
  }
  int lo_peer = h->prm.rank > 0 ? h->prm.rank - 1 : -1;
  int hi_peer = h->prm.rank < h->prm.nranks - 1 ? h->prm.rank + 1 : -1;
  return nccl_halo_exchange_multi(h->comm, lo_peer, hi_peer, lo, nplanes, hi, nplanes, stream);
}

int prepare_eq(wb_fv2d* h) {
  WB_CHECK(exchange_ghost_rows(h, h->weq, 4));
  WB_CUDA(cudaMemsetAsync(&h->ctrl->nonzero_vel, 0, sizeof(int), h->stream));
  dim3 b(128), gr((h->g.nx + 127) / 128, h->g.nyl + 2);
  k_prepare_eq<<<gr, b, 0, h->stream>>>(h->weq, h->eqz, h->g, h->phys, h->ctrl, h->exc, h->eyc);
  WB_LAUNCH_CHECK();
  int nz = 0;
  WB_CUDA(cudaMemcpyAsync(&nz, &h->ctrl->nonzero_vel, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->prm.nranks > 1 && h->comm) {      // every rank must take the same kernels (and the same CFL formula)
    unsigned long long v[2] = {(unsigned long long)(nz & 1), (unsigned long long)((nz >> 1) & 1)};
    unsigned long long* d = reinterpret_cast<unsigned long long*>(h->eqflag);
    WB_CUDA(cudaMemcpyAsync(d, v, sizeof(v), cudaMemcpyHostToDevice, h->stream));
    WB_CHECK(nccl_allreduce_max_u64(h->comm, d, 2, h->stream));
    WB_CUDA(cudaMemcpyAsync(v, d, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
    WB_CUDA(cudaStreamSynchronize(h->stream));
    nz = (int)(v[0] | (v[1] << 1));
  }
  h->fast_ok = (nz & 1) == 0;
  h->eq_analytic = (nz & 2) == 0;
  return WB_OK;
}

bool use_fast(const wb_fv2d* h) { return h->prm.arith == 0 && h->fast_ok; }

// 3-D tensor map (column, row incl. ghosts, plane) of an SoA field for the TMA-fed stage kernel.  The driver entry
// point is fetched through the runtime, so the library keeps no link-time dependency on libcuda.
int make_field_map(const wb_fv2d* h, const double* base, int nplanes, CUtensorMap* out) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    WB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return WB_ERR_CUDA; }
    encode = (encode_fn)fn;
  }
  const Grid& g = h->g;
  const cuuint64_t dims[3] = {(cuuint64_t)g.nx, (cuuint64_t)(g.nyl + 2), (cuuint64_t)nplanes};
  const cuuint64_t strides[2] = {(cuuint64_t)g.pitch * sizeof(double), (cuuint64_t)g.plane * sizeof(double)};
  const cuuint32_t box[3] = {(cuuint32_t)TMA_BOXW, 1u, (cuuint32_t)nplanes};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return WB_ERR_CUDA; }
  return WB_OK;
}

template <int MODE>
int launch_stage(wb_fv2d* h, const double* in, const double* base, double* out, double tend, int max_iter,
                 bool wb_scheme = true, int row_begin = 0, int row_end = -1, cudaStream_t stream = nullptr, bool edge_pair = false) {
  if (row_end < 0) row_end = h->g.nyl;
  if (!stream) stream = h->stream;
  if (row_begin >= row_end) return WB_OK;
  StageArgs A;
  A.in = in; A.base = base; A.out = out; A.weq = h->weq; A.eqz = h->eqz; A.eq_exact = h->eq_analytic ? 0 : 1;
  A.exf = h->exf; A.exc = h->exc; A.eyf = h->eyf; A.eyc = h->eyc;
  A.ctrl = h->ctrl; A.parity = h->parity; A.tend = tend; A.max_iter = max_iter;
  A.row_begin = row_begin; A.row_end = row_end;
  A.pf_rows = h->pf_rows;
  A.rows_cap = 1 << 30;
  A.peer_lo = A.peer_hi = nullptr; A.peer_lo_plane = A.peer_hi_plane = 0;
  {
    const CUtensorMap* mi = (in == h->u) ? &h->map_u : (in == h->w1) ? &h->map_w1 : nullptr;
    const CUtensorMap* mb = (base == h->u) ? &h->map_u : (base == h->w1) ? &h->map_w1 : nullptr;
    const bool tma = use_fast(h) && wb_scheme && h->tma_ok && mi && (MODE != 2 || mb);
    if (edge_pair && !tma) {      // only the TMA kernel knows the paired form
      WB_CHECK((launch_stage<MODE>(h, in, base, out, tend, max_iter, wb_scheme, 0, 1, stream)));
      return launch_stage<MODE>(h, in, base, out, tend, max_iter, wb_scheme, h->g.nyl - 1, h->g.nyl, stream);
    }
  }
  if (use_fast(h) && wb_scheme) {
    const int R = h->march_rows;
    const int ncols = (h->g.nx + MARCH_OUT - 1) / MARCH_OUT;
    dim3 b(MARCH_WARPS * 32), gr((ncols + MARCH_WARPS - 1) / MARCH_WARPS, (A.row_end - A.row_begin + R - 1) / R);
    const CUtensorMap* m_in = (in == h->u) ? &h->map_u : (in == h->w1) ? &h->map_w1 : nullptr;
    const CUtensorMap* m_base = (base == h->u) ? &h->map_u : (base == h->w1) ? &h->map_w1 : nullptr;
    if (h->tma_ok && m_in && (MODE != 2 || m_base)) {
      static bool configured_dev[64] = {};      // function attributes are per device
      bool& configured = configured_dev[h->dev & 63];
      auto kern = k_stage_tma<MODE, MARCH_MIN_BLOCKS>;
      if (!configured) {       // 16 warps x <= 9.3 KiB of ring buffers per SM
        const char* envc = getenv("WB_FV2D_CARVEOUT");
        WB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, envc ? atoi(envc) : (int)cudaSharedmemCarveoutMaxShared));
        WB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_WARPS * tma_warp_bytes(MODE)));
        configured = true;
      }
      int Rt = std::min(R, TMA_MAX_ROWS);            // the per-warp y tables hold one strip
      if ((A.row_end - A.row_begin + Rt - 1) / Rt > 65535) Rt = TMA_MAX_ROWS;      // grid.y limit (ny up to 4.19e6 rows per slab)
      if ((A.row_end - A.row_begin + Rt - 1) / Rt > 65535) { set_error("slab of %d rows is too tall for one launch", A.row_end - A.row_begin); return WB_ERR_ARG; }
      A.rows_cap = Rt;
      if (edge_pair) {
        Rt = h->g.nyl - 1; A.rows_cap = 1; A.row_begin = 0; A.row_end = h->g.nyl;
        if (h->peer.active) {      // (this is the TMA path: stage_with_exchange relies on exactly this condition)
          const wb_fv2d::Peer& Pp = h->peer;
          double* lo = (out == h->u) ? Pp.lo_u : Pp.lo_w1;
          double* hi = (out == h->u) ? Pp.hi_u : Pp.hi_w1;
          if (lo) { A.peer_lo = lo + (size_t)(Pp.lo_nyl + 1) * h->g.pitch; A.peer_lo_plane = Pp.lo_plane; }
          if (hi) { A.peer_hi = hi; A.peer_hi_plane = Pp.hi_plane; }
        }
      }
      const dim3 bt(TMA_WARPS * 32), gt((ncols + TMA_WARPS - 1) / TMA_WARPS, (A.row_end - A.row_begin + Rt - 1) / Rt);
      kern<<<gt, bt, TMA_WARPS * tma_warp_bytes(MODE), stream>>>(*m_in, m_base ? *m_base : *m_in, A, h->g, h->phys, Rt);
    } else {
      k_stage_march<MODE, MARCH_MIN_BLOCKS><<<gr, b, 0, stream>>>(A, h->g, h->phys, R);
    }
  } else {
    dim3 b(128), gr((h->g.nx + 127) / 128, row_end - row_begin);
    if (wb_scheme) k_stage_ref<MODE, true><<<gr, b, 0, stream>>>(A, h->g, h->phys);
    else k_stage_ref<MODE, false><<<gr, b, 0, stream>>>(A, h->g, h->phys);
  }
  WB_LAUNCH_CHECK();
  return WB_OK;
}

// max wave speed of `field` into ctrl->cmax_bits[slot] (all-reduced over ranks)
int launch_max_speed(wb_fv2d* h, const double* field, int slot, bool delta = false) {
  WB_CUDA(cudaMemsetAsync(&h->ctrl->cmax_bits[slot], 0, sizeof(unsigned long long), h->stream));
  dim3 b(256), gr(std::min((h->g.nx + 255) / 256, 64), std::min(h->g.nyl, 592));
  if (h->prm.arith == 0) k_max_speed<true><<<gr, b, 0, h->stream>>>(field, h->g, h->phys, &h->ctrl->cmax_bits[slot], delta ? h->eqz : nullptr);
  else k_max_speed<false><<<gr, b, 0, h->stream>>>(field, h->g, h->phys, &h->ctrl->cmax_bits[slot], nullptr);
  WB_LAUNCH_CHECK();
  if (h->prm.nranks > 1) {
    if (!h->comm) { set_error("nranks > 1 but wb_fv2d_comm_init was not called"); return WB_ERR_STATE; }
    WB_CHECK(nccl_allreduce_max_u64(h->comm, &h->ctrl->cmax_bits[slot], 1, h->stream));
  }
  return WB_OK;
}

int reset_clock(wb_fv2d* h) {
  k_ctrl_reset<<<1, 1, 0, h->stream>>>(h->ctrl);
  WB_LAUNCH_CHECK();
  h->parity = 0;
  return launch_max_speed(h, h->u, 0, use_fast(h));
}

// ---- peer-memory ghost rows ------------------------------------------------------------------------------------
// The boundary-row launch has stored its rows into the neighbours' ghost rows itself; what is left of the "exchange"
// is one flag per direction: k_peer_signal (after the boundary launch, same stream: its stores are complete) publishes
// the exchange number in the neighbours' flag words, k_peer_wait spins until both neighbours have published theirs.
// No WAR hazard needs a second flag: a rank can only write ghost rows of stage s+1 after it has seen the neighbour's
// flag of stage s, and the neighbour raises that flag after the only kernel that reads those ghost rows.
struct PeerRecord {      // what a rank tells the others (all-gathered once at comm_init)
  cudaIpcMemHandle_t u, w1, flags;
  unsigned long long plane;
  int nyl, ok, dev, pad;
};

void peer_close(wb_fv2d* h) {
  wb_fv2d::Peer& P = h->peer;
  void* mapped[] = {P.lo_u, P.lo_w1, P.lo_flags, P.hi_u, P.hi_w1, P.hi_flags};
  for (void* m : mapped)
    if (m) cudaIpcCloseMemHandle(m);
  P.lo_u = P.lo_w1 = P.hi_u = P.hi_w1 = nullptr;
  P.lo_flags = P.hi_flags = nullptr;
  P.active = false;
}

// Collective over the communicator.  Any failure on any rank (no IPC, no peer access, WB_FV2D_P2P=0) leaves every rank on
// the NCCL send/recv path: the decision is all-reduced.
int peer_setup(wb_fv2d* h) {
  wb_fv2d::Peer& P = h->peer;
  const int R = h->prm.nranks, r = h->prm.rank;
  const char* env = getenv("WB_FV2D_P2P");
  int ok = (!env || atoi(env) != 0) && h->tma_ok && h->overlap && h->g.nyl >= 3;
  PeerRecord mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && !P.flags) ok = cudaMalloc(&P.flags, 4 * sizeof(unsigned long long)) == cudaSuccess;
  if (ok) ok = cudaMemsetAsync(P.flags, 0, 4 * sizeof(unsigned long long), h->stream) == cudaSuccess;
  if (ok) ok = cudaIpcGetMemHandle(&mine.u, h->u) == cudaSuccess && cudaIpcGetMemHandle(&mine.w1, h->w1) == cudaSuccess &&
               cudaIpcGetMemHandle(&mine.flags, P.flags) == cudaSuccess;
  cudaGetLastError();
  mine.plane = h->g.plane; mine.nyl = h->g.nyl; mine.ok = ok; mine.dev = h->dev;
  PeerRecord* d_all = nullptr;
  std::vector<PeerRecord> all(R);
  WB_CUDA(cudaMalloc(&d_all, sizeof(PeerRecord) * (R + 1)));
  WB_CUDA(cudaMemcpyAsync(d_all + R, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
  int st = nccl_allgather_bytes(h->comm, d_all + R, d_all, sizeof(PeerRecord), h->stream);
  if (st == WB_OK) {
    WB_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(PeerRecord) * R, cudaMemcpyDeviceToHost, h->stream));
    WB_CUDA(cudaStreamSynchronize(h->stream));
  }
  cudaFree(d_all);
  WB_CHECK(st);
  for (int k = 0; k < R; ++k) ok = ok && all[k].ok;
  auto open = [&](const cudaIpcMemHandle_t& hd, void** out) {
    return cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
  };
  if (ok && r > 0) {
    ok = open(all[r - 1].u, (void**)&P.lo_u) && open(all[r - 1].w1, (void**)&P.lo_w1) && open(all[r - 1].flags, (void**)&P.lo_flags);
    P.lo_plane = all[r - 1].plane; P.lo_nyl = all[r - 1].nyl;
  }
  if (ok && r < R - 1) {
    ok = open(all[r + 1].u, (void**)&P.hi_u) && open(all[r + 1].w1, (void**)&P.hi_w1) && open(all[r + 1].flags, (void**)&P.hi_flags);
    P.hi_plane = all[r + 1].plane; P.hi_nyl = all[r + 1].nyl;
  }
  cudaGetLastError();
  // every rank must have opened its neighbours, or nobody uses the mapping
  unsigned long long* d_bad = h->ctrl ? &h->ctrl->cmax_bits[0] : nullptr;      // scratch word (reset_clock rewrites it)
  unsigned long long bad = ok ? 0ull : 1ull;
  WB_CUDA(cudaMemcpyAsync(d_bad, &bad, sizeof(bad), cudaMemcpyHostToDevice, h->stream));
  WB_CHECK(nccl_allreduce_max_u64(h->comm, d_bad, 1, h->stream));
  WB_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (bad) { peer_close(h); return WB_OK; }
  P.active = true;
  P.seq = 0;
  return WB_OK;
}

// One RK stage in slab mode: the two boundary rows first (on the comm stream, followed by the NCCL send/recv of
// those rows into the neighbours' ghost rows), the interior rows concurrently on the main stream.
template <int MODE>
int stage_with_exchange(wb_fv2d* h, const double* in, const double* base, double* out, double tend) {
  const int nyl = h->g.nyl;
  if (h->prm.nranks <= 1) return launch_stage<MODE>(h, in, base, out, tend, -1);
  if (!h->overlap || nyl < 3) {
    WB_CHECK(launch_stage<MODE>(h, in, base, out, tend, -1));
    return exchange_ghost_rows(h, out, 4);
  }
  WB_CUDA(cudaEventRecord(h->ev_main, h->stream));               // inputs (and their ghosts) are ready
  WB_CUDA(cudaStreamWaitEvent(h->comm_stream, h->ev_main, 0));
  WB_CHECK((launch_stage<MODE>(h, in, base, out, tend, -1, true, 0, nyl, h->comm_stream, true)));     // rows 0 and nyl-1, one launch
  // (the peer stores exist in the TMA-fed kernel only; whether it runs is decided from all-reduced flags: the same on every rank)
  const bool tma_path = use_fast(h) && h->tma_ok && (in == h->u || in == h->w1) && (MODE != 2 || base == h->u || base == h->w1);
  if (h->peer.active && tma_path) {      // the launch stored the rows into the neighbours' ghost rows itself: publish / await the flags
    wb_fv2d::Peer& Pp = h->peer;
    ++Pp.seq;
    // I am the slab ABOVE my lower neighbour (its word [1]) and BELOW my upper one (its word [0])
    WB_CHECK(peer_signal(Pp.lo_flags ? Pp.lo_flags + 1 : nullptr, Pp.hi_flags ? Pp.hi_flags + 0 : nullptr, Pp.seq, h->comm_stream));
    WB_CHECK(peer_wait(Pp.flags, Pp.lo_flags != nullptr, Pp.hi_flags != nullptr, Pp.seq, h->comm_stream));
  } else {
    WB_CHECK(exchange_ghost_rows(h, out, 4, h->comm_stream));
  }
  WB_CUDA(cudaEventRecord(h->ev_comm, h->comm_stream));
  WB_CHECK((launch_stage<MODE>(h, in, base, out, tend, -1, true, 1, nyl - 1, h->stream)));
  WB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_comm, 0));        // next stage needs the ghosts
  return WB_OK;
}

}  // namespace

extern "C" {

int wb_fv2d_create(wb_fv2d** out, const wb_fv2d_params* p) {
  if (!out || !p) { set_error("null argument"); return WB_ERR_ARG; }
  *out = nullptr;
  WB_REQUIRE(p->nvar == 4, "nvar must be 4 (got %d)", p->nvar);
  WB_REQUIRE(p->nx >= 3 && p->ny >= 3, "nx, ny must be >= 3 (got %d x %d)", p->nx, p->ny);
  WB_REQUIRE(p->nx < (1 << 23) && p->ny < (1 << 23), "nx, ny must be < 2^23 (single-precision (i-0.5) of the reference)");
  WB_REQUIRE(p->nequilibrium >= 1 && p->nequilibrium <= 4, "nequilibrium must be 1..4 (got %d)", p->nequilibrium);
  WB_REQUIRE(p->arith == 0 || p->arith == 1, "arith must be 0 (fast) or 1 (reference order)");
  WB_REQUIRE(p->nranks >= 1 && p->rank >= 0 && p->rank < p->nranks, "bad rank/nranks %d/%d", p->rank, p->nranks);
  WB_REQUIRE(p->gamma > 1.0 && p->boxlen_x > 0 && p->boxlen_y > 0 && p->cfl > 0, "gamma>1, boxlen>0, cfl>0 required");
  WB_REQUIRE(p->ny / p->nranks >= 2, "each slab needs at least 2 rows");
  int dev = 0;
  WB_CHECK(select_device(p->device, &dev));
  wb_fv2d* h = new wb_fv2d;
  h->prm = *p;
  h->dev = dev;
  if (const char* e = getenv("WB_FV2D_OVERLAP")) h->overlap = atoi(e);
  if (const char* e = getenv("WB_FV2D_MARCH_ROWS")) h->march_rows = std::max(1, atoi(e));
  if (const char* e = getenv("WB_FV2D_PF_ROWS")) h->pf_rows = std::max(0, atoi(e));
  fill_phys(*p, h->phys);
  Grid& g = h->g;
  g.nx = p->nx; g.ny = p->ny;
  g.j0 = (int)((long long)p->ny * p->rank / p->nranks);
  int j1 = (int)((long long)p->ny * (p->rank + 1) / p->nranks);
  g.nyl = j1 - g.j0;
  g.pitch = (p->nx + 15) / 16 * 16;     // 128-byte aligned rows
  g.plane = (size_t)(g.nyl + 2) * g.pitch;
  auto fail = [&](int s) { wb_fv2d_destroy(h); return s; };
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(WB_ERR_CUDA); }
  h->own_stream = true;
  if (p->nranks > 1) {
    // highest priority: the boundary rows and the NCCL send/recv are enqueued before the interior kernel but would otherwise
    // find every SM's registers taken by it and run at its tail, i.e. not overlapped at all
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming) != cudaSuccess) {
      set_error("comm stream/event creation failed");
      return fail(WB_ERR_CUDA);
    }
  }
  size_t fb = sizeof(double) * 4 * g.plane;
  cudaError_t e;
  if ((e = cudaMalloc(&h->u, fb)) != cudaSuccess || (e = cudaMalloc(&h->w1, fb)) != cudaSuccess ||
      (e = cudaMalloc(&h->weq, fb)) != cudaSuccess || (e = cudaMalloc(&h->eqz, fb / 2)) != cudaSuccess ||
      (e = cudaMalloc(&h->ctrl, sizeof(Ctrl))) != cudaSuccess || (e = cudaMalloc(&h->eqflag, 16)) != cudaSuccess ||
      (e = cudaMallocHost(&h->h_ctrl, sizeof(Ctrl))) != cudaSuccess) {
    set_error("device allocation failed: %s", cudaGetErrorString(e));
    return fail(WB_ERR_CUDA);
  }
  cudaMemsetAsync(h->u, 0, fb, h->stream);
  cudaMemsetAsync(h->w1, 0, fb, h->stream);
  cudaMemsetAsync(h->weq, 0, fb, h->stream);
  cudaMemsetAsync(h->eqz, 0, fb / 2, h->stream);
  cudaMemsetAsync(h->ctrl, 0, sizeof(Ctrl), h->stream);
  // separable equilibrium tables: exp(-a*xf(i)), exp(-a*xc(i)), exp(-a*yf(j)), exp(-a*yc(j)) (local rows)
  {
    const Phys& P = h->phys;
    size_t nxf = g.nx + 1, nxc = g.nx, nyf = g.nyl + 1, nyc = g.nyl;
    std::vector<double> t(nxf + nxc + nyf + nyc);
    double z = (P.neq == 4) ? 0.0 : 1.0;
    for (size_t i = 0; i < nxf; ++i) t[i] = z * std::exp(-P.a * ((double)i * P.dx));
    for (size_t i = 0; i < nxc; ++i) t[nxf + i] = z * std::exp(-P.a * ((double)((float)(i + 1) - 0.5f) * P.dx));
    for (size_t j = 0; j < nyf; ++j) t[nxf + nxc + j] = z * std::exp(-P.a * ((double)(g.j0 + j) * P.dx /* sic :513 */));
    for (size_t j = 0; j < nyc; ++j) t[nxf + nxc + nyf + j] = z * std::exp(-P.a * ((double)((float)(g.j0 + j + 1) - 0.5f) * P.dy));
    if ((e = cudaMalloc(&h->tab, sizeof(double) * t.size())) != cudaSuccess) { set_error("table allocation failed"); return fail(WB_ERR_CUDA); }
    if ((e = cudaMemcpy(h->tab, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice)) != cudaSuccess) {
      set_error("table upload failed: %s", cudaGetErrorString(e)); return fail(WB_ERR_CUDA);
    }
    h->exf = h->tab; h->exc = h->tab + nxf; h->eyf = h->tab + nxf + nxc; h->eyc = h->tab + nxf + nxc + nyf;
  }
  if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) { set_error("init failed: %s", cudaGetErrorString(e)); return fail(WB_ERR_CUDA); }
  // tensor maps for the TMA-fed stage kernel (the box is 34 columns wide; narrower grids use the LDG march)
  {
    const char* env = getenv("WB_FV2D_TMA");
    if (p->arith == 0 && g.nx >= TMA_BOXW && !(env && atoi(env) == 0)) {
      int st = make_field_map(h, h->u, 4, &h->map_u);
      if (st == WB_OK) st = make_field_map(h, h->w1, 4, &h->map_w1);
      if (st != WB_OK) return fail(st);
      h->tma_ok = true;
    }
  }
  *out = h;
  return WB_OK;
}

int wb_fv2d_destroy(wb_fv2d* h) {
  if (!h) return WB_OK;
  cudaSetDevice(h->dev);
  if (h->stream) cudaStreamSynchronize(h->stream);
  output_wait(&h->out_job);
  if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
  if (h->peer.active && h->comm) {      // nobody frees a buffer a neighbour still has mapped: close, meet, then free
    peer_close(h);
    nccl_allreduce_max_u64(h->comm, &h->ctrl->cmax_bits[0], 1, h->stream);
    cudaStreamSynchronize(h->stream);
  }
  cudaFree(h->peer.flags);
  nccl_comm_destroy(h->comm);
  cudaFree(h->u); cudaFree(h->w1); cudaFree(h->weq); cudaFree(h->eqz); cudaFree(h->stage); cudaFree(h->tab);
  cudaFree(h->ctrl); cudaFree(h->eqflag);
  if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  if (h->comm_stream) { cudaStreamSynchronize(h->comm_stream); cudaStreamDestroy(h->comm_stream); }
  if (h->ev_main) cudaEventDestroy(h->ev_main);
  if (h->ev_comm) cudaEventDestroy(h->ev_comm);
  delete h;
  return WB_OK;
}

int wb_fv2d_local_rows(const wb_fv2d* h, int* j0, int* nrows) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  if (j0) *j0 = h->g.j0;
  if (nrows) *nrows = h->g.nyl;
  return WB_OK;
}

int wb_fv2d_set_stream(wb_fv2d* h, void* cuda_stream) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  h->stream = (cudaStream_t)cuda_stream;
  return WB_OK;
}

int wb_fv2d_comm_init(wb_fv2d* h, const void* id128) {
  if (!h || !id128) { set_error("null argument"); return WB_ERR_ARG; }
  WB_REQUIRE(h->prm.nranks > 1, "comm_init needs nranks > 1");
  WB_CUDA(cudaSetDevice(h->dev));
  if (h->comm) { peer_close(h); nccl_comm_destroy(h->comm); h->comm = nullptr; }
  WB_CHECK(nccl_comm_create(&h->comm, id128, h->prm.rank, h->prm.nranks));
  return peer_setup(h);
}

const char* wb_fv2d_exchange_kind(const wb_fv2d* h) {
  if (!h || h->prm.nranks <= 1) return "none";
  return h->peer.active ? "p2p" : "nccl";
}

int wb_fv2d_upload(wb_fv2d* h, const double* u, const double* w_eq) {
  if (!h || !u || !w_eq) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  WB_CHECK(h2d_state(h, w_eq, h->weq));
  WB_CHECK(prepare_eq(h));
  WB_CHECK(h2d_state(h, u, h->u, use_fast(h)));        // fused kernels: resident state in delta form
  WB_CHECK(exchange_ghost_rows(h, h->u, 4));
  WB_CHECK(reset_clock(h));
  h->resident = true;
  return WB_OK;
}

int wb_fv2d_init_device(wb_fv2d* h, int ninit, double eta) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  WB_REQUIRE(ninit >= 1 && ninit <= 4, "ninit must be 1..4 (got %d)", ninit);
  WB_CUDA(cudaSetDevice(h->dev));
  dim3 b(128), gr((h->g.nx + 127) / 128, h->g.nyl + 2);
  k_init<<<gr, b, 0, h->stream>>>(h->u, h->weq, h->g, h->phys, ninit, eta, 1, 1);
  WB_LAUNCH_CHECK();
  WB_CHECK(prepare_eq(h));
  if (use_fast(h)) {
    k_to_delta<<<gr, b, 0, h->stream>>>(h->u, h->eqz, h->g);
    WB_LAUNCH_CHECK();
  }
  WB_CHECK(reset_clock(h));
  h->resident = true;
  return WB_OK;
}

int wb_fv2d_get_initial_conditions(wb_fv2d* h, int ninit, double eta, double* u_out, double* w_eq_out) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  WB_REQUIRE(ninit >= 1 && ninit <= 4, "ninit must be 1..4 (got %d)", ninit);
  WB_CUDA(cudaSetDevice(h->dev));
  // uses w1 / stage as scratch so that a resident state is not disturbed
  double* scratch_u = h->w1;
  double* scratch_w = nullptr;
  WB_CUDA(cudaMalloc(&scratch_w, sizeof(double) * 4 * h->g.plane));
  dim3 b(128), gr((h->g.nx + 127) / 128, h->g.nyl + 2);
  k_init<<<gr, b, 0, h->stream>>>(scratch_u, scratch_w, h->g, h->phys, ninit, eta, 1, 1);
  wb::g_launches.fetch_add(1);
  int st = WB_OK;
  if (cudaGetLastError() != cudaSuccess) { set_error("k_init launch failed"); st = WB_ERR_CUDA; }
  if (st == WB_OK && u_out) st = d2h_state(h, scratch_u, u_out);
  if (st == WB_OK && w_eq_out) st = d2h_state(h, scratch_w, w_eq_out);
  cudaStreamSynchronize(h->stream);
  cudaFree(scratch_w);
  return st;
}

int wb_fv2d_reset_clock(wb_fv2d* h) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  if (!h->resident) { set_error("no resident state (call upload or init_device first)"); return WB_ERR_STATE; }
  WB_CUDA(cudaSetDevice(h->dev));
  return reset_clock(h);
}

int wb_fv2d_step_async(wb_fv2d* h, int nsteps, double tend) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  if (!h->resident) { set_error("no resident state (call upload or init_device first)"); return WB_ERR_STATE; }
  WB_CUDA(cudaSetDevice(h->dev));
  for (int s = 0; s < nsteps; ++s) {
    // stage 1: w1 = u + dt*L(u)                                   benchmark_2d.f90:246-247
    WB_CHECK(stage_with_exchange<1>(h, h->u, nullptr, h->w1, tend));
    // stage 2: u = .5u + .5w1 + .5dt*L(w1), fused max speed        benchmark_2d.f90:249-250, :241
    WB_CHECK(stage_with_exchange<2>(h, h->w1, h->u, h->u, tend));
    if (h->prm.nranks > 1) WB_CHECK(nccl_allreduce_max_u64(h->comm, &h->ctrl->cmax_bits[h->parity ^ 1], 1, h->stream));
    h->parity ^= 1;
  }
  return WB_OK;
}

int wb_fv2d_sync(wb_fv2d* h, int* iters_out, double* t_out, double* last_dt_out, double* last_cmax_out) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  WB_CUDA(cudaMemcpyAsync(h->h_ctrl, h->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->peer.active) {      // a ghost-row wait that gave up (a neighbour never published its flag) must not pass silently
    unsigned long long err = 0;
    WB_CUDA(cudaMemcpy(&err, h->peer.flags + 2, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) { set_error("peer-memory ghost-row exchange %llu timed out (a neighbouring rank stopped)", err); return WB_ERR_NCCL; }
  }
  int s = (h->h_ctrl->iter[1] > h->h_ctrl->iter[0]) ? 1 : 0;
  if (iters_out) *iters_out = h->h_ctrl->iter[s];
  if (t_out) *t_out = h->h_ctrl->t[s];
  if (last_dt_out) *last_dt_out = h->h_ctrl->dt_last;
  if (last_cmax_out) *last_cmax_out = h->h_ctrl->cmax_last;
  return WB_OK;
}

int wb_fv2d_download(wb_fv2d* h, double* u_out) {
  if (!h || !u_out) { set_error("null argument"); return WB_ERR_ARG; }
  if (!h->resident) { set_error("no resident state"); return WB_ERR_STATE; }
  WB_CUDA(cudaSetDevice(h->dev));
  return d2h_state(h, h->u, u_out, use_fast(h));
}

int wb_fv2d_evolve(wb_fv2d* h, double* u_inout, const double* w_eq, double tend, int max_iter, int* iters_out,
                   double* t_out, double* last_dt_out) {
  if (!h || !u_inout || !w_eq) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CHECK(wb_fv2d_upload(h, u_inout, w_eq));
  int iters = 0;
  double t = 0.0, dt = 0.0;
  const int batch = 16;
  for (;;) {
    if (!(t < tend) || (max_iter >= 0 && iters >= max_iter)) break;
    int n = batch;
    if (max_iter >= 0 && max_iter - iters < n) n = max_iter - iters;
    WB_CHECK(wb_fv2d_step_async(h, n, tend));
    WB_CHECK(wb_fv2d_sync(h, &iters, &t, &dt, nullptr));
  }
  WB_CHECK(wb_fv2d_download(h, u_inout));
  if (iters_out) *iters_out = iters;
  if (t_out) *t_out = t;
  if (last_dt_out) *last_dt_out = dt;
  return WB_OK;
}

int wb_fv2d_output_file(wb_fv2d* h, const char* path) {
  if (!h || !path) { set_error("null argument"); return WB_ERR_ARG; }
  if (!h->resident) { set_error("no resident state"); return WB_ERR_STATE; }
  WB_REQUIRE(h->prm.nranks == 1, "output_file: single-GPU handles only (a slab holds part of the table)");
  WB_CUDA(cudaSetDevice(h->dev));
  double* tab = nullptr;
  WB_CUDA(cudaMalloc(&tab, sizeof(double) * 3 * (size_t)h->g.nx * h->g.nyl));
  dim3 b(128), gr((h->g.nx + 127) / 128, h->g.nyl);
  k_pack_output<<<gr, b, 0, h->stream>>>(h->u, use_fast(h) ? h->eqz : nullptr, h->g, h->phys, tab);
  WB_LAUNCH_CHECK();
  return output_start(&h->out_job, h->dev, h->stream, tab, (size_t)h->g.nx * h->g.nyl, 3, path);
}

int wb_fv2d_output_wait(wb_fv2d* h) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  return output_wait(&h->out_job);
}

static int update_common(wb_fv2d* h, const double* u, const double* w_eq, double* dudt, bool wb_scheme) {
  if (!h || !u || !w_eq || !dudt) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  // stateless: uses the resident buffers as scratch -> invalidates a resident state
  h->resident = false;
  WB_CHECK(h2d_state(h, w_eq, h->weq));
  WB_CHECK(prepare_eq(h));
  WB_CHECK(h2d_state(h, u, h->u, use_fast(h) && wb_scheme));
  WB_CHECK(exchange_ghost_rows(h, h->u, 4));
  WB_CHECK(launch_stage<0>(h, h->u, nullptr, h->w1, 0.0, -1, wb_scheme));
  return d2h_state(h, h->w1, dudt);
}

int wb_fv2d_compute_update_exact(wb_fv2d* h, const double* u, const double* w_eq, double* dudt) {
  return update_common(h, u, w_eq, dudt, true);
}
int wb_fv2d_compute_update(wb_fv2d* h, const double* u, const double* w_eq, double* dudt) {
  return update_common(h, u, w_eq, dudt, false);
}

int wb_fv2d_compute_max_speed(wb_fv2d* h, const double* u, double* cmax) {
  if (!h || !u || !cmax) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  h->resident = false;
  WB_CHECK(h2d_state(h, u, h->w1));
  WB_CHECK(launch_max_speed(h, h->w1, 0));
  unsigned long long bits = 0;
  WB_CUDA(cudaMemcpyAsync(&bits, &h->ctrl->cmax_bits[0], sizeof(bits), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  memcpy(cmax, &bits, sizeof(double));
  return WB_OK;
}

}  // extern "C"
