// Shared device-side definitions of the 2D DG path: grid / physics / control structs, the plane addressing macro, the
// pointwise physics of 2d/benchmark_2d_dg.f90 in the reference's operation order, the neighbour rules.
// Included by dg2d.cu and dg2d_split.cu.
#pragma once
#include "common.cuh"
#include "dg_basis.h"

namespace wb { namespace dg {

struct DgGrid {
  int nx, ny, m, nm;      // elements (ny = LOCAL rows incl. the two ghost rows in slab mode), order, modes per element (m*m)
  size_t ne;              // nx*ny: elements per plane as allocated
  // slab decomposition along y (one process per GPU).  Ghost rows are ordinary rows of the local arrays: every kernel
  // treats them like any other row (their results are garbage and are overwritten by the exchange), only the
  // neighbour rule, the wave-speed scan and the host <-> device copies know about them.
  int nyg;                // global ny (the reference wraps x-face neighbours with it, 2d/benchmark_2d_dg.f90:1338)
  int j0;                 // global row index of local row 0 (0 without slabs, first owned row - 1 with slabs)
  int slab;               // 1: y neighbours are jc +- 1 clamped to the local array; 0: boundary condition applied in y
  size_t e_off, ne_own;   // first owned element (nx in slab mode) and number of owned elements
};
struct DgPhys {
  double gamma, gm1a, gm1b;   // gamma, gamma-1.0 (real(4) literal) and gamma-1. -- the same value, kept apart for clarity
  double oneoverdx, dx;
  double eps, M;
  double dt_num;              // cfl*min(1/9, gll_w_1/2)
  int bc, source, flux_id, ninit;
  double rho_floor, p_floor;  // (double)10e-10f of compute_primitive (:898) and the 1d-10 of the sound speed (:980)
  int gsep, grow;             // gravity field found separable at upload (gx = f(column, qx), gy = f(row, qy), e.g. the shipped
                              // grad_phi_case 1 on the tensor-product grid): the stage kernel then reads gx from the one row
                              // `grow` and gy from column 0 -- L2-resident lines instead of 144 B of DRAM per element and stage
};
struct DgCtrl {
  double t, dt, tend;
  int iter, max_iter, skip;
  double cs_max, vx, vy, speed_max;
  long long kstar;
};

#define PL(base, g, v, mode) ((base) + ((size_t)(v) * (g).nm + (mode)) * (g).ne)

// ------------------------------------------------------------------------------------ pointwise physics
// compute_primitive :891-902 (density floored at the real(4) literal 10e-10)
__device__ __forceinline__ void prim(const DgPhys& P, const double u[4], double w[4]) {
  w[0] = fmax(u[0], (double)10e-10f);
  w[1] = u[1] / w[0];
  w[2] = u[2] / w[0];
  w[3] = P.gm1a * (u[3] - 0.5 * w[0] * (w[1] * w[1] + w[2] * w[2]));
}
// compute_conservative :905-917
__device__ __forceinline__ void cons(const DgPhys& P, const double w[4], double u[4]) {
  u[0] = w[0];
  u[1] = w[0] * w[1];
  u[2] = w[0] * w[2];
  u[3] = w[3] / P.gm1b + 0.5 * (w[0] * (w[1] * w[1] + w[2] * w[2]));
}
// compute_flux :919-944 (volume nodes; flux(1) uses the floored density)
__device__ __forceinline__ void flux_nodes(const DgPhys& P, const double u[4], double f1[4], double f2[4]) {
  double w[4];
  prim(P, u, w);
  f2[0] = w[0] * w[2];
  f2[1] = w[0] * w[1] * w[2];
  f2[2] = w[2] * u[2] + w[3];
  f2[3] = w[2] * u[3] + w[2] * w[3];
  f1[0] = w[0] * w[1];
  f1[1] = w[1] * u[1] + w[3];
  f1[2] = w[0] * w[1] * w[2];
  f1[3] = w[1] * u[3] + w[1] * w[3];
}
// compute_flux_int :946-965, one direction
template <int DIR>
__device__ __forceinline__ void flux_int(const DgPhys& P, const double u[4], double f[4]) {
  double w[4];
  prim(P, u, w);
  if (DIR == 1) {
    f[0] = w[1] * u[0];
    f[1] = w[1] * u[1] + w[3];
    f[2] = w[0] * w[1] * w[2];
    f[3] = w[1] * u[3] + w[1] * w[3];
  } else {
    f[0] = w[2] * u[0];
    f[1] = w[0] * w[1] * w[2];
    f[2] = w[2] * u[2] + w[3];
    f[3] = w[2] * u[3] + w[2] * w[3];
  }
}
// compute_speed :872-889
__device__ __forceinline__ void speed(const DgPhys& P, const double u[4], double& cs, double& vx, double& vy, double& spd) {
  double w[4];
  prim(P, u, w);
  cs = sqrt(P.gamma * fmax(w[3], 1e-10) / fmax(w[0], 1e-10));
  vx = w[1];
  vy = w[2];
  spd = sqrt(w[1] * w[1] + w[2] * w[2]) + cs;
}
// compute_hllflux :1008-1026 ('hll2'): isotropic speeds |v| +- cs, the direction only enters through fl, fr
__device__ __forceinline__ void hllflux(const DgPhys& P, const double ul[4], const double ur[4], const double fl[4], const double fr[4],
                                        double fh[4]) {
  double csl, csr, vxl, vyl, vxr, vyr, sl, sr;
  speed(P, ul, csl, vxl, vyl, sl);
  speed(P, ur, csr, vxr, vyr, sr);
  const double ml = sqrt(vxl * vxl + vyl * vyl), mr = sqrt(vxr * vxr + vyr * vyr);
  const double a_plus = fmax(0.0, fmax(csl + ml, csr + mr));
  const double a_minus = fmax(0.0, fmax(-(csl - ml), -(csr - mr)));
#pragma unroll
  for (int v = 0; v < 4; ++v) fh[v] = (a_plus * fl[v] + a_minus * fr[v] - a_plus * a_minus * (ur[v] - ul[v])) / (a_plus + a_minus);
}
// compute_hllcflux :1030-1134 ('hllc') as shipped: misplaced parenthesis in the right star energy (:1075, :1116), wleft(2)
// in the right star state of the y branch (:1114), fluxes of compute_flux (:919-944); untouched output when no branch fires
template <int DIR>
__device__ __forceinline__ void hllcflux(const DgPhys& P, const double ul[4], const double ur[4], double fh[4]) {
  double wl[4], wr[4], csl, csr, vxl, vyl, vxr, vyr, sl, sr, f1[4], f2[4], usl[4], usr[4];
  prim(P, ul, wl);
  prim(P, ur, wr);
  speed(P, ul, csl, vxl, vyl, sl);
  speed(P, ur, csr, vxr, vyr, sr);
  constexpr int n = (DIR == 1) ? 1 : 2, t = (DIR == 1) ? 2 : 1;
  const double v_l = (DIR == 1) ? vxl : vyl, v_r = (DIR == 1) ? vxr : vyr;
  const double SL = fmin(v_l, v_r) - fmax(csl, csr), SR = fmax(v_l, v_r) + fmax(csl, csr);
  const double SM = (wr[0] * v_r * (SR - v_r) - wl[0] * v_l * (SL - v_l) + wl[3] - wr[3]) / (wr[0] * (SR - v_r) - wl[0] * (SL - v_l));
  usl[0] = ul[0] * (SL - v_l) / (SL - SM);
  usl[n] = usl[0] * SM;
  usl[t] = usl[0] * wl[t];
  usl[3] = usl[0] * (ul[3] / ul[0] + (SM - wl[n]) * (SM + wl[3] / (wl[0] * (SL - wl[n]))));
  usr[0] = ur[0] * (SR - v_r) / (SR - SM);
  usr[n] = usr[0] * SM;
  usr[t] = usr[0] * ((DIR == 1) ? wr[t] : wl[t]);
  usr[3] = usr[0] * (ur[3] / ur[0] + (SM - wr[n] * (SM + wr[3] / (wr[0] * (SR - wr[n])))));
  if (SL > 0.0) {
    flux_nodes(P, ul, f1, f2);
#pragma unroll
    for (int v = 0; v < 4; ++v) fh[v] = (DIR == 1) ? f1[v] : f2[v];
  } else if (SL <= 0 && SM > 0) {
    flux_nodes(P, ul, f1, f2);
#pragma unroll
    for (int v = 0; v < 4; ++v) fh[v] = ((DIR == 1) ? f1[v] : f2[v]) + SL * (usl[v] - ul[v]);
  } else if (SR >= 0 && SM <= 0) {
    flux_nodes(P, ur, f1, f2);
#pragma unroll
    for (int v = 0; v < 4; ++v) fh[v] = ((DIR == 1) ? f1[v] : f2[v]) + SR * (usr[v] - ur[v]);
  } else if (SR < 0) {
    flux_nodes(P, ur, f1, f2);
#pragma unroll
    for (int v = 0; v < 4; ++v) fh[v] = (DIR == 1) ? f1[v] : f2[v];
  }
}
// compute_num_flux :991-1006 -> compute_llflux :968-988 | compute_hllflux | compute_hllcflux; flux_id 0 ('llf', the shipped
// value, matches no branch) leaves the flux at its initial 0
template <int DIR>
__device__ __forceinline__ void num_flux(const DgPhys& P, const double ul[4], const double ur[4], double nf[4]) {
  nf[0] = nf[1] = nf[2] = nf[3] = 0.0;
  if (P.flux_id == 0) return;
  if (P.flux_id == 3) { hllcflux<DIR>(P, ul, ur, nf); return; }
  double fl[4], fr[4], csl, csr, vxl, vyl, vxr, vyr, sl, sr;
  flux_int<DIR>(P, ul, fl);
  flux_int<DIR>(P, ur, fr);
  if (P.flux_id == 2) { hllflux(P, ul, ur, fl, fr, nf); return; }
  speed(P, ul, csl, vxl, vyl, sl);
  speed(P, ur, csr, vxr, vyr, sr);
  double cmax = (DIR == 1) ? fmax(fabs(vxr + csr), fabs(vxl + csl)) : fmax(fabs(vyr + csr), fabs(vyl + csl));
#pragma unroll
  for (int v = 0; v < 4; ++v) nf[v] = 0.5 * (fr[v] + fl[v]) + 0.5 * cmax * (ul[v] - ur[v]);
}

// get_boundary_conditions :777-824 on a 0-based index that may be -1 or n
__device__ __forceinline__ int bc_index(int bc, int idx, int n) {
  if (bc == 1) { if (idx < 0) idx = n - 1; else if (idx >= n) idx = 0; }
  else if (bc == 2 || bc == 3) { if (idx < 0) idx = 0; else if (idx >= n) idx = n - 1; }
  return idx;
}

// row of the y neighbour: boundary condition on a whole grid, plain +-1 (ghost rows) on a slab
__device__ __forceinline__ int y_nb(const DgGrid& g, int bc, int jc) {
  return g.slab ? min(max(jc, 0), g.ny - 1) : bc_index(bc, jc, g.ny);
}

template <int M>
__device__ __forceinline__ void load_modes(const double* __restrict__ u, const DgGrid& g, size_t e, double d[4][M][M]) {
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int i = 0; i < M; ++i) d[v][i][j] = PL(u, g, v, j * M + i)[e];
}

// edge traces :1253-1314.  SIDE 0 left (xi=-1), 1 right (xi=+1): points along y; 2 bottom, 3 top: points along x.
template <int M, int SIDE>
__device__ __forceinline__ void trace(const double d[4][M][M], const Basis& B, double out[M][4]) {
#pragma unroll
  for (int q = 0; q < M; ++q)
#pragma unroll
    for (int v = 0; v < 4; ++v) out[q][v] = 0.0;
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int q = 0; q < M; ++q)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          if (SIDE == 0) out[q][v] = out[q][v] + d[v][i][j] * B.Em[i] * B.P[q][j];
          if (SIDE == 1) out[q][v] = out[q][v] + d[v][i][j] * B.Ep[i] * B.P[q][j];
          if (SIDE == 2) out[q][v] = out[q][v] + d[v][i][j] * B.Em[j] * B.P[q][i];
          if (SIDE == 3) out[q][v] = out[q][v] + d[v][i][j] * B.Ep[j] * B.P[q][i];
        }
}

// solve_for_t :312-362
__device__ __forceinline__ double solve_for_t(const DgPhys& P, const double u[4], const double ua[4]) {
  const double eps = P.eps;
  double pa = ua[0], mxa = ua[1], mya = ua[2], ea = ua[3];
  double pj = u[0], mxj = u[1], myj = u[2], ej = u[3];
  double a = 2.0 * (pj - pa) * (ej - ea) - (mxj - mxa) * (mxj - mxa) - (myj - mya) * (myj - mya);
  double b = 2.0 * (pj - pa) * (ea - eps / (P.gamma - 1)) + 2.0 * pa * (ej - ea) - 2.0 * (mxa * (mxj - mxa) + mya * (myj - mya));
  double c = 2.0 * pa * ea - (mxa * mxa + mya * mya) - 2.0 * eps * pa / P.gm1a;
  b = b / a;
  c = c / a;
  double D = sqrt(fabs(b * b - 4 * c));
  double t1 = 0.5 * (-b - D), t2 = 0.5 * (-b + D), t;
  if ((t1 > -eps) && (t1 < (double)1.0f + eps)) t = t1;
  else if ((t2 > -eps) && (t2 < (double)1.0f + eps)) t = t2;
  else t = 0.0;
  t = fmin(1.0, t);
  t = fmax(0.0, t);
  return t;
}

}}  // namespace wb::dg
