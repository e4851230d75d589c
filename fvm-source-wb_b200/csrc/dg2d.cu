// 2D modal DG (tensor Legendre basis, Gauss-Legendre quadrature) RK-stage kernels + C-ABI.
// Replaces 2d/benchmark_2d_dg.f90: evolve :624-775, compute_max_speed :826-870, compute_update :1137-1479,
// apply_limiter :1516-1555, the transforms :497-592, and the limiters of 2d/limiters.f90 it dispatches to
// ('ONP' :478-654, 'HIO' :1441-1583, '1OR' :203-309, 'LOW' :769-860).
//
// Device layout: structure-of-arrays planes  plane(v, m)[jc*nx + ic],  m = jm*M + im, i.e.
// [var][mode][row][column] with the element column contiguous -> one thread per element reads/writes
// every plane coalesced.
//
// Arithmetic: this first version keeps the REFERENCE'S OPERATION ORDER everywhere (file compiled with
// -fmad=false, IEEE div/sqrt): accumulation order of every quadrature sum, left-to-right products, the
// real(4)-rounded RK coefficients.  With the basis tables computed by the same recurrences on the host the
// results agree with the CPU restatement bit for bit.
#include "common.cuh"
#include "dg_basis.h"
#include "dg2d_common.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

namespace wb { namespace dg {

// ------------------------------------------------------------------------------------ layout kernels
// host u(nvar,nx,ny,mx,my) == [mode][jc][ic][4]  <->  device planes
__global__ void k_dg_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, DgGrid g) {
  size_t eh = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // host arrays hold the owned rows only
  int mode = blockIdx.y;
  if (eh >= g.ne_own) return;
  const size_t e = eh + g.e_off;
  const double2* src = reinterpret_cast<const double2*>(aos + ((size_t)mode * g.ne_own + eh) * 4);
  double2 a = src[0], b = src[1];
  PL(soa, g, 0, mode)[e] = a.x; PL(soa, g, 1, mode)[e] = a.y; PL(soa, g, 2, mode)[e] = b.x; PL(soa, g, 3, mode)[e] = b.y;
}
__global__ void k_dg_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, DgGrid g) {
  size_t eh = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int mode = blockIdx.y;
  if (eh >= g.ne_own) return;
  const size_t e = eh + g.e_off;
  double2* dst = reinterpret_cast<double2*>(aos + ((size_t)mode * g.ne_own + eh) * 4);
  dst[0] = make_double2(PL(soa, g, 0, mode)[e], PL(soa, g, 1, mode)[e]);
  dst[1] = make_double2(PL(soa, g, 2, mode)[e], PL(soa, g, 3, mode)[e]);
}

// ------------------------------------------------------------------------------------ transforms :497-592
template <int M>
__global__ void k_nodes_from_modes(const double* __restrict__ modes, double* __restrict__ nodes, DgGrid g, Basis B) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne) return;
  double d[4][M][M];
  load_modes<M>(modes, g, e, d);
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int i = 0; i < M; ++i) {
        double a = 0.0;
#pragma unroll
        for (int in = 0; in < M; ++in)
#pragma unroll
          for (int jn = 0; jn < M; ++jn) a = a + d[v][in][jn] * B.P[i][in] * B.P[j][jn];
        PL(nodes, g, v, j * M + i)[e] = a;
      }
}
template <int M>
__device__ __forceinline__ void modes_from_nodes_el(const double nd[4][M][M], const Basis& B, double md[4][M][M]) {
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double a = 0.0;
#pragma unroll
        for (int xq = 0; xq < M; ++xq)
#pragma unroll
          for (int yq = 0; yq < M; ++yq) a = a + 0.25 * nd[v][xq][yq] * B.P[xq][i] * B.P[yq][j] * B.wq[xq] * B.wq[yq];
        md[v][i][j] = a;
      }
}
template <int M>
__device__ __forceinline__ void nodes_from_modes_el(const double md[4][M][M], const Basis& B, double nd[4][M][M]) {
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double a = 0.0;
#pragma unroll
        for (int in = 0; in < M; ++in)
#pragma unroll
          for (int jn = 0; jn < M; ++jn) a = a + md[v][in][jn] * B.P[i][in] * B.P[j][jn];
        nd[v][i][j] = a;
      }
}
template <int M>
__global__ void k_modes_from_nodes(const double* __restrict__ nodes, double* __restrict__ modes, DgGrid g, Basis B) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne) return;
  double nd[4][M][M], md[4][M][M];
  load_modes<M>(nodes, g, e, nd);
  modes_from_nodes_el<M>(nd, B, md);
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int i = 0; i < M; ++i) PL(modes, g, v, j * M + i)[e] = md[v][i][j];
}

// Is the gravity field separable?  gx(qy,qx; j,i) == gx(0,qx; j0,i) and gy(qy,qx; j,i) == gy(qy,0; j,0) for every owned
// element and node, compared bit for bit (so a NaN says no).  flag[0] is cleared on the first difference.
__global__ void k_grad_sep_check(const double* __restrict__ gx, const double* __restrict__ gy, DgGrid g, int row0, int row1,
                                 int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = row0 + blockIdx.y;
  if (i >= g.nx || j >= row1) return;
  bool same = true;
  for (int qy = 0; qy < g.m; ++qy)
    for (int qx = 0; qx < g.m; ++qx) {
      const size_t e = (size_t)(qy * g.m + qx) * g.ne + (size_t)j * g.nx + i;
      same = same && gx[e] == gx[(size_t)qx * g.ne + (size_t)row0 * g.nx + i] &&
             gy[e] == gy[(size_t)(qy * g.m) * g.ne + (size_t)j * g.nx];
    }
  if (!same) *flag = 0;
}

// grad_phi :1599-1644 at every node, once per upload (x, y are static)
__global__ void k_grad_phi(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ gx,
                           double* __restrict__ gy, size_t n, int grad_phi_case) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  if (grad_phi_case == 1) { gx[k] = x[k]; gy[k] = y[k]; return; }
  const double epsilon = 0.25, delta_r = (double)0.1f, x_center = 3., y_center = 3.;
  double x_dash = x[k] - x_center, y_dash = y[k] - y_center;
  double r = sqrt(x_dash * x_dash + y_dash * y_dash);
  if (r > 0.5 - 0.5 * delta_r) {
    gx[k] = -(x_dash) / ((r * r) * r);
    gy[k] = -(y_dash) / ((r * r) * r);
  } else {
    gx[k] = -(x_dash) / (r * (r * r + epsilon * epsilon));
    gy[k] = -(y_dash) / (r * (r * r + epsilon * epsilon));
  }
}
// special_boundary_conditions :1481-1514 (ninit == 12): freeze flags per (mode index, element)
__global__ void k_freeze_mask(const double* __restrict__ x, const double* __restrict__ y, unsigned char* __restrict__ fz,
                              size_t n, double xc, double yc) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double xd = x[k] - xc, yd = y[k] - yc;
  fz[k] = (sqrt(xd * xd + yd * yd) > 2.0) ? 1 : 0;
}

// get_coords :93-120 and get_initial_conditions :122-466 (cases 1 and 2) on the device, for grids whose nodal arrays
// the host cannot reasonably hold.  Case 1 needs the global minimum of the nodal density (w(4) = minval(w(1)), :141):
// pass 0 computes it (atomicMin on the bit pattern of a positive double), pass 1 writes the state.
__device__ __forceinline__ void node_xy(const DgGrid& g, const DgPhys& P, const Basis& B, size_t e, int mode, double dy,
                                         double& x, double& y) {
  const int ic = (int)(e % g.nx), jc = (int)(e / g.nx);
  const int ni = mode % g.m, nj = mode / g.m;
  x = (double)((float)(ic + 1) - 0.5f) * P.dx + P.dx / 2.0 * B.xq[ni];
  y = (double)((float)(g.j0 + jc + 1) - 0.5f) * dy + dy / 2.0 * B.xq[nj];
}
// Keplerian velocity profile of the rotating disks (cases 7 and 11, :290-306, :405-420); the reference's second branch can
// never fire (its condition is contained in the first one)
__device__ __forceinline__ void disk_velocity(double x_dash, double y_dash, double r, double delta_r, double& w2, double& w3) {
  const double r32 = pow(r, 1.5);
  if (r <= (double)0.5f - delta_r) { w2 = 0.; w3 = 0.; }
  else if ((r > (double)0.5f - delta_r) && (r <= 2 + delta_r)) { w2 = -(y_dash / r32); w3 = x_dash / r32; }
  else if ((r > 2 + delta_r) && (r <= 2 + 2 * delta_r)) {
    w2 = y_dash / r32 / delta_r * (r - (2 + delta_r)) - y_dash / r32;
    w3 = -(x_dash / r32 / delta_r * (r - (2 + delta_r))) + x_dash / r32;
  } else { w2 = 0.; w3 = 0.; }
}
// primitive initial state at one node, get_initial_conditions :122-466, all twelve cases; returns true when w[3] is the
// global minimum of w[0] (cases 1, 10, 11: filled in by the caller's second pass)
__device__ __forceinline__ bool dg_ic_prim(int ninit, double x, double y, const DgPhys& P, double boxlen_x, double boxlen_y, double eta,
                                           double w[4]) {
  const double dpi = 3.141592653589793;       // acos(-1d0)
  w[0] = 0.; w[1] = 0.; w[2] = 0.; w[3] = 0.;
  switch (ninit) {
    case 1: {
      const double ax = x - boxlen_x / 2., ay = y - boxlen_y / 2.;
      w[0] = exp(-((ax * ax + ay * ay) * 10));
      w[1] = 1.0; w[2] = 1.0;
      return true;
    }
    case 2: {
      const double rho_0 = (double)1.21f, p_0 = 1., gg = 1.;
      const double ee = exp(-(rho_0 * gg / p_0) * (x + y));
      const double bx = x - (double)0.3f, by = y - (double)0.3f;
      w[0] = rho_0 * ee;
      w[3] = p_0 * ee + eta * exp(-(100 * (rho_0 * gg / p_0) * (bx * bx + by * by)));
      return false;
    }
    case 3:
      if (x >= 0.5 && y >= 0.5) { w[0] = 1.; w[3] = 1.; }
      else if (x < 0.5 && y >= 0.5) { w[0] = (double)0.5197f; w[1] = (double)-0.7259f; w[3] = (double)0.4f; }
      else if (x < 0.5 && y < 0.5) { w[0] = (double)0.1072f; w[1] = (double)-0.7259f; w[2] = (double)-1.4045f; w[3] = (double)0.0439f; }
      else { w[0] = (double)0.2579f; w[2] = (double)-1.4045f; w[3] = (double)0.15f; }
      return false;
    case 4:
      if (x >= 0.5 && y >= 0.5) { w[0] = 1.5; w[3] = 1.5; }
      else if (x < 0.5 && y >= 0.5) { w[0] = (double)0.5323f; w[1] = (double)1.206f; w[3] = (double)0.3f; }
      else if (x < 0.5 && y < 0.5) { w[0] = (double)0.138f; w[1] = (double)1.206f; w[2] = (double)1.206f; w[3] = (double)0.029f; }
      else { w[0] = (double)0.5323f; w[2] = (double)1.206f; w[3] = (double)0.3f; }
      return false;
    case 5:
      if (x + y >= 0.5) { w[0] = 1.; w[3] = 1.; }
      else { w[0] = 0.125; w[3] = (double)0.4f; }
      return false;
    case 6: {      // isentropic vortex
      const double r2 = (x - 5) * (x - 5) + (y - 5) * (y - 5);
      w[0] = 1. * pow(1. - (P.gamma - 1.) * 5 / (8 * P.gamma * (dpi * dpi)) * exp(1 - r2), 1 / (P.gamma - 1));
      w[1] = 2 + 5. / (2 * dpi) * exp(-1 - r2 / 2.) * (-y + 5.);
      w[2] = 2 + 5. / (2 * dpi) * exp(-1 - r2 / 2.) * (x - 5.);
      w[3] = pow(w[0], P.gamma);
      return false;
    }
    case 7: {      // smooth rotating disk (boxlen 6 x 6)
      const double p_0 = (double)10e-5f, rho_0 = (double)10e-5f, rho_d = 1., delta_r = (double)0.1f;
      const double x_dash = x - 3., y_dash = y - 3.;
      const double r = sqrt(x_dash * x_dash + y_dash * y_dash);
      w[3] = p_0;
      if (r < (double)0.5f - delta_r / 2.) w[0] = rho_0;
      else if ((r < (double)0.5f + delta_r / 2.) && (r > (double)0.5f - delta_r / 2.))
        w[0] = (rho_d - rho_0) / delta_r * (r - ((double)0.5f - delta_r / 2.)) + rho_0;
      else if ((r >= (double)0.5f + delta_r / 2.) && (r <= 2 - delta_r / 2.)) w[0] = rho_d;
      else if ((r > 2 - delta_r / 2.) && (r < 2 + delta_r / 2.)) w[0] = (rho_0 - rho_d) / delta_r * (r - (2 - delta_r / 2.)) + rho_d;
      else if (r >= 2 + delta_r / 2.) w[0] = rho_0;
      disk_velocity(x_dash, y_dash, r, delta_r, w[1], w[2]);
      return false;
    }
    case 8: {      // square advection
      const double x_dash = x - 0.5, y_dash = y - 0.5;
      w[0] = ((fabs(x_dash) <= 0.25) && (fabs(y_dash) <= 0.25)) ? 4.0 : 1.0;
      w[2] = 10.0; w[3] = 1.0;
      return false;
    }
    case 9: {      // 1-d discontinuous pulse advection
      w[0] = (fabs(y - 0.5) <= 0.25) ? 4. : 1.;
      w[2] = 1.0; w[3] = 1.;
      return false;
    }
    case 10: {     // Gaussian density, w(4) = minval(w(1))
      const double rho_0 = (double)1.21f, p_0 = 1., gg = 1.;
      const double ax = x - boxlen_x / 2., ay = y - boxlen_y / 2.;
      w[0] = rho_0 * exp(-(rho_0 * gg / p_0) * (ax * ax + ay * ay) * 20);
      w[1] = 1.0; w[2] = 1.0;
      return true;
    }
    case 11: {     // 100 % smooth rotating disk, w(4) = minval(w(1))
      const double delta_r = (double)0.1f;
      const double x_dash = x - 3., y_dash = y - 3.;
      const double r = sqrt(x_dash * x_dash + y_dash * y_dash);
      const double e = exp(-2 * ((r - 2.) * (r - 2.)));
      w[0] = e * e;
      disk_velocity(x_dash, y_dash, r, delta_r, w[1], w[2]);
      return true;
    }
    default: {     // case 12: Keplerian disk, softened potential
      const double rho_d = 1.0, GM = 1., H = (double)0.05f, epsilon = 0.25;
      const double x_dash = x - 0.5 * boxlen_x, y_dash = y - 0.5 * boxlen_y;
      const double r = sqrt(x_dash * x_dash + y_dash * y_dash);
      const double q = sqrt(r * r + epsilon * epsilon);
      const double cs_m = H * sqrt(GM / q);
      w[0] = rho_d;
      w[3] = cs_m * cs_m * rho_d;
      w[1] = -(y_dash * sqrt(GM / q - H * H * GM / q));
      w[2] = x_dash * sqrt(GM / q - H * H * GM / q);
      return false;
    }
  }
}
// pass 0 (cases whose pressure is the global minimum of the nodal density, w(4) = minval(w(1)) :141, :377, :428): that minimum
// (atomicMin on the bit pattern of a positive double); pass 1 writes the conservative nodal state
__global__ void k_dg_init(double* __restrict__ nodes, double* __restrict__ xy, DgGrid g, DgPhys P, Basis B, int ninit, double eta,
                          double boxlen_x, double boxlen_y, unsigned long long* minbits, int pass, double shift_x = 0.0,
                          double shift_y = 0.0) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int mode = blockIdx.y;
  if (e >= g.ne) return;
  const double dy = boxlen_y / (double)g.nyg;
  double x, y;
  node_xy(g, P, B, e, mode, dy, x, y);
  if (shift_x != 0.0 || shift_y != 0.0) {      // the initial state translated by (shift_x, shift_y) in the periodic box
    x = x - shift_x; y = y - shift_y;
    x = x - boxlen_x * floor(x / boxlen_x); y = y - boxlen_y * floor(y / boxlen_y);
  }
  double w[4];
  const bool needs_min = dg_ic_prim(ninit, x, y, P, boxlen_x, boxlen_y, eta, w);
  if (pass == 0) {      // minimum over the owned rows only (ghost rows lie outside the box or belong to a neighbour)
    if (needs_min && e >= g.e_off && e < g.e_off + g.ne_own) atomicMin(minbits, (unsigned long long)__double_as_longlong(w[0]));
    return;
  }
  if (needs_min) w[3] = __longlong_as_double((long long)*minbits);
  double u[4];
  cons(P, w, u);
#pragma unroll
  for (int v = 0; v < 4; ++v) PL(nodes, g, v, mode)[e] = u[v];
  if (xy) { xy[(size_t)mode * g.ne + e] = x; xy[((size_t)g.nm + mode) * g.ne + e] = y; }
}

// output_file (2d/benchmark_2d_dg.f90:468-495): one row per element, icell outer / jcell inner: x, y of node (1,1) and
// w - w_eq for the variables var..nvar there (compute_primitive of the nodal value; get_equilibrium_solution :594-622)
template <int M>
__global__ void k_dg_pack_output(const double* __restrict__ modes, DgGrid g, DgPhys P, Basis B, double boxlen_y, int var, int nequilibrium,
                                 double* __restrict__ tab) {
  size_t eo = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (eo >= g.ne_own) return;
  const size_t e = eo + g.e_off;
  double d[4][M][M], u[4], w[4];
  load_modes<M>(modes, g, e, d);
#pragma unroll
  for (int v = 0; v < 4; ++v) {            // node (1,1) as get_nodes_from_modes evaluates it (:575-587)
    double a = 0.0;
#pragma unroll
    for (int in = 0; in < M; ++in)
#pragma unroll
      for (int jn = 0; jn < M; ++jn) a = a + d[v][in][jn] * B.P[0][in] * B.P[0][jn];
    u[v] = a;
  }
  prim(P, u, w);
  double x, y;
  node_xy(g, P, B, e, 0, boxlen_y / (double)g.nyg, x, y);
  double weq[4] = {0.0, 0.0, 0.0, 0.0};
  if (nequilibrium == 1) { weq[0] = exp(-(x + y)); weq[3] = exp(-(x + y)); }
  else if (nequilibrium == 2) {
    const double rho_0 = (double)1.21f, p_0 = 1., gg = 1.;
    weq[0] = rho_0 * exp(-(rho_0 * gg / p_0) * (x + y)); weq[3] = p_0 * exp(-(rho_0 * gg / p_0) * (x + y));
  }
  const int ic = (int)(e % g.nx), jc = (int)(e / g.nx) - (g.slab ? 1 : 0);
  const int ncol = 2 + (4 - var + 1);
  double* row = tab + ((size_t)ic * (g.ne_own / g.nx) + jc) * ncol;
  row[0] = x; row[1] = y;
  for (int v = var; v <= 4; ++v) row[2 + v - var] = w[v - 1] - weq[v - 1];
}

// ------------------------------------------------------------------------------------ compute_update :1137-1479
template <int M>
__global__ void __launch_bounds__(128) k_dg_update(const double* __restrict__ du, const double* __restrict__ gx,
                                                   const double* __restrict__ gy, const unsigned char* __restrict__ fz,
                                                   double* __restrict__ dudt, DgGrid g, DgPhys P, Basis B,
                                                   const DgCtrl* __restrict__ ctrl) {
  if (ctrl && ctrl->skip) return;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne) return;
  const int ic = (int)(e % g.nx), jc = (int)(e / g.nx);
  double d[4][M][M];
  load_modes<M>(du, g, e, d);
  // nodal values and fluxes at the volume quadrature points :1203-1204
  double f1[4][M][M], f2[4][M][M], s[4][M][M];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double uq[4], a1[4], a2[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        double a = 0.0;
#pragma unroll
        for (int in = 0; in < M; ++in)
#pragma unroll
          for (int jn = 0; jn < M; ++jn) a = a + d[v][in][jn] * B.P[i][in] * B.P[j][jn];
        uq[v] = a;
      }
      flux_nodes(P, uq, a1, a2);
#pragma unroll
      for (int v = 0; v < 4; ++v) { f1[v][i][j] = a1[v]; f2[v][i][j] = a2[v]; }
      // source at the node :1414-1424
      if (P.source == 2) {
        double w[4];
        prim(P, uq, w);
        double g1 = gx[(size_t)(j * M + i) * g.ne + e], g2 = gy[(size_t)(j * M + i) * g.ne + e];
        s[0][i][j] = 0.;
        s[1][i][j] = w[0] * g1;
        s[2][i][j] = w[0] * g2;
        s[3][i][j] = w[0] * (w[1] * g1 + w[2] * g2);
      } else if (P.source == 3) {
        s[0][i][j] = -1.0 * uq[0]; s[1][i][j] = 0.0; s[2][i][j] = 0.0; s[3][i][j] = 0.0;
      } else {
        s[0][i][j] = 0.0; s[1][i][j] = 0.0; s[2][i][j] = 0.0; s[3][i][j] = 0.0;
      }
    }
  // own traces and the facing traces of the four neighbours :1253-1314; x-face neighbours are wrapped with ny
  // (get_boundary_conditions(.,2) in the x sweep, :1338-1339; nx == ny is enforced at create)
  double tl[M][4], tr[M][4], tb[M][4], tt[M][4];      // own left/right/bottom/top
  trace<M, 0>(d, B, tl); trace<M, 1>(d, B, tr); trace<M, 2>(d, B, tb); trace<M, 3>(d, B, tt);
  double FL[M][4], FR[M][4], GB[M][4], GT[M][4];
  {
    double nb[4][M][M], tn[M][4];
    const int il = bc_index(P.bc, ic - 1, g.nyg), ir = bc_index(P.bc, ic + 1, g.nyg);
    const int jb = y_nb(g, P.bc, jc - 1), jt = y_nb(g, P.bc, jc + 1);
    load_modes<M>(du, g, (size_t)jc * g.nx + il, nb);
    trace<M, 1>(nb, B, tn);                            // left neighbour's right trace
#pragma unroll
    for (int q = 0; q < M; ++q) num_flux<1>(P, tn[q], tl[q], FL[q]);
    load_modes<M>(du, g, (size_t)jc * g.nx + ir, nb);
    trace<M, 0>(nb, B, tn);                            // right neighbour's left trace
#pragma unroll
    for (int q = 0; q < M; ++q) num_flux<1>(P, tr[q], tn[q], FR[q]);
    load_modes<M>(du, g, (size_t)jb * g.nx + ic, nb);
    trace<M, 3>(nb, B, tn);                            // bottom neighbour's top trace
#pragma unroll
    for (int q = 0; q < M; ++q) num_flux<2>(P, tn[q], tb[q], GB[q]);
    load_modes<M>(du, g, (size_t)jt * g.nx + ic, nb);
    trace<M, 2>(nb, B, tn);                            // top neighbour's bottom trace
#pragma unroll
    for (int q = 0; q < M; ++q) num_flux<2>(P, tt[q], tn[q], GT[q]);
  }
  // volume, edge and source integrals per mode, then the update :1207-1244, :1372-1411, :1427-1466
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        double vol1 = 0.0, vol2 = 0.0, sv = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0, e4 = 0.0;
#pragma unroll
        for (int in = 0; in < M; ++in)
#pragma unroll
          for (int jn = 0; jn < M; ++jn) vol1 = vol1 + f1[v][in][jn] * B.dP[in][i] * B.wq[in] * B.P[jn][j] * B.wq[jn];
#pragma unroll
        for (int in = 0; in < M; ++in)
#pragma unroll
          for (int jn = 0; jn < M; ++jn) vol2 = vol2 + f2[v][in][jn] * B.dP[jn][j] * B.wq[jn] * B.P[in][i] * B.wq[in];
#pragma unroll
        for (int q = 0; q < M; ++q) {
          e1 = e1 + FR[q][v] * B.Ep[i] * B.P[q][j] * B.wq[q];
          e2 = e2 + FL[q][v] * B.Em[i] * B.P[q][j] * B.wq[q];
        }
#pragma unroll
        for (int q = 0; q < M; ++q) {
          e3 = e3 + GT[q][v] * B.Ep[j] * B.P[q][i] * B.wq[q];
          e4 = e4 + GB[q][v] * B.Em[j] * B.P[q][i] * B.wq[q];
        }
#pragma unroll
        for (int in = 0; in < M; ++in)
#pragma unroll
          for (int jn = 0; jn < M; ++jn) sv = sv + s[v][in][jn] * B.P[in][i] * B.wq[in] * B.P[jn][j] * B.wq[jn];
        double r = (P.oneoverdx * vol1 + P.oneoverdx * vol2 - P.oneoverdx * (e1 - e2) - P.oneoverdx * (e3 - e4)) / 2. + sv / 4.;
        if (fz && fz[(size_t)(j * M + i) * g.ne + e]) r = 0.0;
        PL(dudt, g, v, j * M + i)[e] = r;
      }
}

// ------------------------------------------------------------------------------------ RK combinations :672-747
// out = c0*A0 [+ c1*A1 [+ c2*A2 + c3*A3]] + (cd*dt)*D, evaluated left to right as the Fortran array expressions.
// c0 == 1 with NA == 1 reproduces `delta_u + c*dt*dudt` (1.0*x == x exactly).
template <int NA>
__global__ void k_dg_axpy(double* __restrict__ out, const double* __restrict__ A0, double c0, const double* __restrict__ A1,
                          double c1, const double* __restrict__ A2, double c2, const double* __restrict__ A3, double c3,
                          const double* __restrict__ D, double cd, size_t n, const DgCtrl* __restrict__ ctrl) {
  if (ctrl->skip) return;
  const double cdt = cd * ctrl->dt;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    if (NA == 0) { out[k] = A0[k]; continue; }
    double r = (NA == 1 && c0 == 1.0) ? A0[k] : c0 * A0[k];
    if (NA >= 2) r = r + c1 * A1[k];
    if (NA >= 4) { r = r + c2 * A2[k]; r = r + c3 * A3[k]; }
    out[k] = r + cdt * D[k];
  }
}

// ------------------------------------------------------------------------------------ limiters (2d/limiters.f90)
__device__ __forceinline__ double sign1(double x) { return copysign(1.0, x); }
__device__ __forceinline__ double minmod(double x, double y, double z) {                       // :18-28
  double s = sign1(x);
  if (sign1(y) == s && sign1(z) == s) return s * fmin(fmin(fabs(x), fabs(y)), fabs(z));
  return 0.0;
}
__device__ __forceinline__ double generalized_minmod(const DgPhys& P, double x, double y, double z) {  // :30-54
  if (fabs(x) < P.M * (P.dx * P.dx)) return x;
  return minmod(x, y, z);
}
__device__ __forceinline__ double minmod2d(double u, double dlx, double dly, double drx, double dry) {  // :57-78
  double s = sign1(u);
  if (sign1(dlx) == s && sign1(dly) == s && sign1(drx) == s && sign1(dry) == s)
    return s * fmin(fmin(fmin(fmin(fabs(u), fabs(dly)), fabs(dlx)), fabs(dry)), fabs(drx));
  return 0.0;
}
// compute_set :438-475: point (q, r) of the "left" family (GLL in x, GL in y) and of the "right" family
template <int M>
__device__ __forceinline__ void set_point(const double el[4][M][M], const Basis& B, int q, int r, double ul[4], double ur[4]) {
#pragma unroll
  for (int v = 0; v < 4; ++v) { ul[v] = 0.0; ur[v] = 0.0; }
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        ul[v] = ul[v] + el[v][i][j] * B.Pg[r][i] * B.P[q][j];
        ur[v] = ur[v] + el[v][i][j] * B.P[q][i] * B.Pg[r][j];
      }
}
// compute_positivity ('ONP') :478-654, one element
template <int M>
__device__ __forceinline__ void positivity_el(const DgPhys& P, const Basis& B, double el[4][M][M]) {
  if (M == 1) return;
  const double uavg[4] = {el[0][0][0], el[1][0][0], el[2][0][0], el[3][0][0]};
  // 1. density: theta from the minimum over the point set (the set is stored left family first, then right)
  double p_min = 0.0;
  bool first = true;
  for (int fam = 0; fam < 2; ++fam)
    for (int q = 0; q < M; ++q)
      for (int r = 0; r < B.gll; ++r) {
        double ul[4], ur[4];
        set_point<M>(el, B, q, r, ul, ur);
        double val = fam == 0 ? ul[0] : ur[0];
        p_min = first ? val : fmin(p_min, val);
        first = false;
      }
  const double theta = fmin(fabs((uavg[0] - P.eps) / (uavg[0] - p_min)), 1.0);
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j)
      if (i != 0 || j != 0) el[0][i][j] = theta * el[0][i][j];
  // 2. pressure
  double t_min = 1.;
  for (int fam = 0; fam < 2; ++fam)
    for (int q = 0; q < M; ++q)
      for (int r = 0; r < B.gll; ++r) {
        double ul[4], ur[4], w[4], t;
        set_point<M>(el, B, q, r, ul, ur);
        const double* pt = fam == 0 ? ul : ur;
        prim(P, pt, w);
        if (w[3] > P.eps) t = 1.;
        else t = solve_for_t(P, pt, uavg);
        if (t_min >= t) t_min = t;
      }
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j)
      if (i != 0 || j != 0) {
#pragma unroll
        for (int v = 0; v < 4; ++v) el[v][i][j] = t_min * el[v][i][j];
      }
}
template <int M>
__device__ __forceinline__ void store_modes(double* __restrict__ u, const DgGrid& g, size_t e, const double d[4][M][M]) {
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int i = 0; i < M; ++i) PL(u, g, v, j * M + i)[e] = d[v][i][j];
}
template <int M>
__global__ void k_limiter_onp(double* __restrict__ u, DgGrid g, DgPhys P, Basis B, const DgCtrl* __restrict__ ctrl) {
  if (ctrl && ctrl->skip) return;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne) return;
  double el[4][M][M];
  load_modes<M>(u, g, e, el);
  positivity_el<M>(P, B, el);
  store_modes<M>(u, g, e, el);
}
// limiter_low_order ('LOW') :769-860
template <int M>
__global__ void k_limiter_low(double* __restrict__ u, DgGrid g, const DgCtrl* __restrict__ ctrl) {
  if (ctrl && ctrl->skip) return;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne || M == 1) return;
  for (int v = 0; v < 4; ++v) {
    for (int j = 1; j < M; ++j) PL(u, g, v, j * M + 0)[e] = 0.0;
    for (int i = 1; i < M; ++i) PL(u, g, v, 0 * M + i)[e] = 0.0;
  }
}
// high_order_limiter ('HIO') :1478-1583 without its trailing compute_positivity (launched separately).
// Reads the un-limited modes of the element and of its 4 neighbours from `u`, writes `un`.
template <int M>
__global__ void k_limiter_hio(const double* __restrict__ u, double* __restrict__ un, DgGrid g, DgPhys P,
                              const DgCtrl* __restrict__ ctrl) {
  if (ctrl && ctrl->skip) return;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne) return;
  const int ic = (int)(e % g.nx), jc = (int)(e / g.nx);
  const size_t eL = (size_t)jc * g.nx + bc_index(P.bc, ic - 1, g.nx), eR = (size_t)jc * g.nx + bc_index(P.bc, ic + 1, g.nx);
  const size_t eB = (size_t)y_nb(g, P.bc, jc - 1) * g.nx + ic, eT = (size_t)y_nb(g, P.bc, jc + 1) * g.nx + ic;
  // mode (a,b), 1-based as in the reference -> plane index (b-1)*M + (a-1)
#define MD(el, a, b) (PL(u, g, v, ((b) - 1) * M + ((a) - 1))[el])
  for (int v = 0; v < 4; ++v) {
    // u_new = u for this element/variable
    for (int m = 0; m < M * M; ++m) PL(un, g, v, m)[e] = PL(u, g, v, m)[e];
    auto limiting = [&](int a, int b) {                                              // :1441-1476
      double coeff_j = (2.0 * (double)(a - 1) + 1.0) * (2 * (double)(b - 1) - 1);
      double coeff_i = (2.0 * (double)(b - 1) + 1.0) * (2 * (double)(a - 1) - 1);
      double coeff_u = (2.0 * (double)(a - 1) + 1.0) * (2.0 * (double)(b - 1) + 1.0);
      double central_u = MD(e, a, b);
      double d_r_y = (MD(eT, a, b - 1) - MD(e, a, b - 1)) * coeff_j;
      double d_l_y = (MD(e, a, b - 1) - MD(eB, a, b - 1)) * coeff_j;
      double d_r_x = (MD(eR, a - 1, b) - MD(e, a - 1, b)) * coeff_i;
      double d_l_x = (MD(e, a - 1, b) - MD(eL, a - 1, b)) * coeff_i;
      return minmod2d(central_u * coeff_u, d_r_y, d_l_y, d_r_x, d_l_x) / coeff_u;
    };
    int done = 0;
    for (int a = M; a >= 2; --a) {
      double limited = limiting(a, a);
      if (limited != MD(e, a, a)) PL(un, g, v, (a - 1) * M + (a - 1))[e] = limited;
      else break;
      for (int b = a - 1; b >= 2; --b) {
        double l1 = limiting(a, b), l2 = limiting(b, a);
        if ((fabs(l1 - MD(e, a, b)) < P.eps) && (fabs(l2 - MD(e, b, a)) < P.eps)) { done = 1; break; }
        PL(un, g, v, (b - 1) * M + (a - 1))[e] = l1;
        PL(un, g, v, (a - 1) * M + (b - 1))[e] = l2;
      }
      if (done == 1) break;
      double coeff_y = (2 * (double)(a - 1) + 1), coeff_u = (2 * (double)(a - 1) + 1);
      double d_r_y = MD(eT, a - 1, 1) - MD(e, a - 1, 1);
      double d_l_y = MD(e, a - 1, 1) - MD(eB, a - 1, 1);
      double d_r_x = MD(eR, 1, a - 1) - MD(e, 1, a - 1);
      double d_l_x = MD(e, 1, a - 1) - MD(eL, 1, a - 1);
      double l1 = generalized_minmod(P, MD(e, 1, a) * coeff_u, d_r_y * coeff_y, d_l_y * coeff_y) / coeff_u;
      double l2 = generalized_minmod(P, MD(e, a, 1) * coeff_u, d_r_x * coeff_y, d_l_x * coeff_y) / coeff_u;
      if ((l1 == MD(e, 1, a)) && (l2 == MD(e, a, 1))) break;
      PL(un, g, v, (a - 1) * M + 0)[e] = l1;      // u_new(1,a)
      PL(un, g, v, 0 * M + (a - 1))[e] = l2;      // u_new(a,1)
    }
  }
#undef MD
}
// limiter_positivity_2 ('PO3', 2d/limiters.f90:1587-1711) in two element-parallel passes, reference operation order.
// get_matrix_decomp (2d/benchmark_2d_dg.f90:2096-2177) with k = (0, 1); rows as in the reference's literals.
__device__ __forceinline__ void po3_matrices(const DgPhys& P, const double wa[4], double lev[4][4], double rev[4][4]) {
  const double kappa = P.gamma - 1;
  const double k1 = (double)0.f, k2 = (double)1.f;
  const double ca = sqrt(P.gamma * wa[3] / wa[0]);
  const double phis = sqrt(1 / 2.f * kappa * (wa[1] * wa[1] + wa[2] * wa[2]));
  const double beta = 1.f / (2 * (ca * ca));
  const double theta = k1 * wa[1] + k2 * wa[2];
  rev[0][0] = 1 - phis * phis / (ca * ca); rev[0][1] = kappa * wa[1] / (ca * ca); rev[0][2] = kappa * wa[2] / (ca * ca); rev[0][3] = -kappa / (ca * ca);
  rev[1][0] = -(k2 * wa[1] - k1 * wa[2]); rev[1][1] = k2; rev[1][2] = -k1; rev[1][3] = 0.0;
  rev[2][0] = beta * (phis * phis - ca * theta); rev[2][1] = beta * (k1 * ca - kappa * wa[1]); rev[2][2] = beta * (k2 * ca - kappa * wa[2]); rev[2][3] = beta * kappa;
  rev[3][0] = beta * (phis * phis + ca * theta); rev[3][1] = -beta * (k1 * ca + kappa * wa[1]); rev[3][2] = -beta * (k2 * ca + kappa * wa[2]); rev[3][3] = beta * kappa;
  lev[0][0] = 1.0; lev[0][1] = 0.0; lev[0][2] = 1.0; lev[0][3] = 1.0;
  lev[1][0] = wa[1]; lev[1][1] = k2; lev[1][2] = wa[1] + k1 * ca; lev[1][3] = wa[1] - k1 * ca;
  lev[2][0] = wa[2]; lev[2][1] = -k1; lev[2][2] = wa[2] + k2 * ca; lev[2][3] = wa[2] - k2 * ca;
  lev[3][0] = phis * phis / (kappa); lev[3][1] = k2 * wa[1] - k1 * wa[2];
  lev[3][2] = (phis * phis + ca * ca) / kappa + ca * theta; lev[3][3] = (phis * phis + ca * ca) / kappa - ca * theta;
}
// nodal values of the element, and the (1,1) mode of its modal PRIMITIVE variables (compute_characteristics :2031-2040)
template <int M>
__device__ __forceinline__ void po3_prepare(const DgPhys& P, const Basis& B, const double md[4][M][M], double nd[4][M][M], double wa[4]) {
  nodes_from_modes_el<M>(md, B, nd);
  double wn[4][M][M];
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double uu[4] = {nd[0][i][j], nd[1][i][j], nd[2][i][j], nd[3][i][j]}, ww[4];
      prim(P, uu, ww);
#pragma unroll
      for (int v = 0; v < 4; ++v) wn[v][i][j] = ww[v];
    }
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    double a = 0.0;
#pragma unroll
    for (int xq = 0; xq < M; ++xq)
#pragma unroll
      for (int yq = 0; yq < M; ++yq) a = a + 0.25 * wn[v][xq][yq] * B.P[xq][0] * B.P[yq][0] * B.wq[xq] * B.wq[yq];
    wa[v] = a;
  }
}
// pass A: modes of the characteristic variables lev * u_n of every element -> cm
template <int M>
__global__ void __launch_bounds__(128) k_limiter_po3_a(const double* __restrict__ u, double* __restrict__ cm, DgGrid g, DgPhys P, Basis B,
                                                       const DgCtrl* __restrict__ ctrl) {
  if (ctrl && ctrl->skip) return;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne || M == 1) return;
  double md[4][M][M], nd[4][M][M], wa[4], lev[4][4], rev[4][4];
  load_modes<M>(u, g, e, md);
  po3_prepare<M>(P, B, md, nd, wa);
  po3_matrices(P, wa, lev, rev);
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const double un[4] = {nd[0][i][j], nd[1][i][j], nd[2][i][j], nd[3][i][j]};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) s = s + lev[r][k] * un[k];
        nd[r][i][j] = s;
      }
    }
  modes_from_nodes_el<M>(nd, B, md);
  store_modes<M>(cm, g, e, md);
}
// pass B: minmod of the linear characteristic modes against the PERIODIC neighbours' means (the reference wraps with nx in
// both passes whatever bc is), back through rev, nodal reset of density / pressure below 1d-10, projection -> out
template <int M>
__global__ void __launch_bounds__(128) k_limiter_po3_b(const double* u, const double* __restrict__ cm, double* out,      /* u may be out (in place) */
                                                       DgGrid g, DgPhys P, Basis B, const DgCtrl* __restrict__ ctrl) {
  if (ctrl && ctrl->skip) return;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne || M == 1) return;
  const int ic = (int)(e % g.nx), jc = (int)(e / g.nx);
  const size_t eL = (size_t)jc * g.nx + (ic == 0 ? g.nx - 1 : ic - 1), eR = (size_t)jc * g.nx + (ic == g.nx - 1 ? 0 : ic + 1);
  // y neighbours: a slab finds them in its ghost rows (ring of ranks), a whole grid wraps
  const int jb = g.slab ? max(jc - 1, 0) : (jc == 0 ? g.ny - 1 : jc - 1), jt = g.slab ? min(jc + 1, g.ny - 1) : (jc == g.ny - 1 ? 0 : jc + 1);
  const size_t eB = (size_t)jb * g.nx + ic, eT = (size_t)jt * g.nx + ic;
  double ul[4][M][M];
  load_modes<M>(cm, g, e, ul);
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const double u_center = PL(cm, g, v, 0)[e];
    {      // x pass: mode (2,1)
      const double u_left = PL(cm, g, v, 0)[eL], u_right = PL(cm, g, v, 0)[eR], u_deriv = PL(cm, g, v, 1)[e];
      const double l = minmod(u_deriv, (u_center - u_left), (u_right - u_center));
      const bool drop = fabs(l - u_deriv) > (double)0.01f * fabs(u_deriv);
      ul[v][1][0] = l;
      if (drop) {
#pragma unroll
        for (int i = 1; i < M; ++i) ul[v][i][0] = 0.0;
        ul[v][M - 1][M - 1] = 0.0;
      }
    }
    {      // y pass: mode (1,2)
      const double u_left = PL(cm, g, v, 0)[eB], u_right = PL(cm, g, v, 0)[eT], u_deriv = PL(cm, g, v, M)[e];
      const double l = minmod(u_deriv, (u_center - u_left), (u_right - u_center));
      const bool drop = fabs(l - u_deriv) > (double)0.01f * fabs(u_deriv);
      ul[v][0][1] = l;
      if (drop) {
#pragma unroll
        for (int j = 1; j < M; ++j) ul[v][0][j] = 0.0;
        ul[v][M - 1][M - 1] = 0.0;
      }
    }
  }
  double md[4][M][M], nd[4][M][M], ch[4][M][M], wa[4], lev[4][4], rev[4][4];
  load_modes<M>(u, g, e, md);
  po3_prepare<M>(P, B, md, nd, wa);
  po3_matrices(P, wa, lev, rev);
  nodes_from_modes_el<M>(ul, B, ch);
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const double c[4] = {ch[0][i][j], ch[1][i][j], ch[2][i][j], ch[3][i][j]};
      double uc[4], ww[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) s = s + rev[r][k] * c[k];
        uc[r] = s;
      }
      prim(P, uc, ww);
      if (ww[0] < 1e-10) ww[0] = (double)1e-5f;
      if (ww[3] < 1e-10) ww[3] = (double)1e-5f;
      cons(P, ww, uc);
#pragma unroll
      for (int r = 0; r < 4; ++r) nd[r][i][j] = uc[r];
    }
  modes_from_nodes_el<M>(nd, B, md);
  store_modes<M>(out, g, e, md);
}

// ---- the neighbour-reading limiters in the FUSED flow (arith 0): the stage kernel writes the un-limited stage result to
// a scratch field `u`; these kernels read it (element + 4 neighbours) and write the limited result to `out` in ONE pass, with
// the reference's operation order (same device code as the unfused kernels above: same bits).  `out2` is the second result
// of SSPRK(5,4) stage 4 (w5, :700-704): the stage kernel left it without its k3*w4 term, which needs the LIMITED w4.
// A skipped step (ctrl->skip) hands its input through unchanged.
template <int M>
__global__ void __launch_bounds__(128) k_limiter_hio_onp(const double* __restrict__ u, double* __restrict__ out, double* __restrict__ out2,
                                                         double k3, DgGrid g, DgPhys P, Basis B, const DgCtrl* __restrict__ ctrl) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne) return;
  double un[4][M][M];
  load_modes<M>(u, g, e, un);
  if (ctrl->skip) { store_modes<M>(out, g, e, un); return; }
  const int ic = (int)(e % g.nx), jc = (int)(e / g.nx);
  const size_t eL = (size_t)jc * g.nx + bc_index(P.bc, ic - 1, g.nx), eR = (size_t)jc * g.nx + bc_index(P.bc, ic + 1, g.nx);
  const size_t eB = (size_t)y_nb(g, P.bc, jc - 1) * g.nx + ic, eT = (size_t)y_nb(g, P.bc, jc + 1) * g.nx + ic;
  if (M > 1) {
#define MD(el, a, b) (PL(u, g, v, ((b) - 1) * M + ((a) - 1))[el])
#define UN(a, b) un[v][(a) - 1][(b) - 1]
    for (int v = 0; v < 4; ++v) {              // high_order_limiter :1478-1583, u_new in registers
      auto limiting = [&](int a, int b) {      // :1441-1476
        double coeff_j = (2.0 * (double)(a - 1) + 1.0) * (2 * (double)(b - 1) - 1);
        double coeff_i = (2.0 * (double)(b - 1) + 1.0) * (2 * (double)(a - 1) - 1);
        double coeff_u = (2.0 * (double)(a - 1) + 1.0) * (2.0 * (double)(b - 1) + 1.0);
        double central_u = MD(e, a, b);
        double d_r_y = (MD(eT, a, b - 1) - MD(e, a, b - 1)) * coeff_j;
        double d_l_y = (MD(e, a, b - 1) - MD(eB, a, b - 1)) * coeff_j;
        double d_r_x = (MD(eR, a - 1, b) - MD(e, a - 1, b)) * coeff_i;
        double d_l_x = (MD(e, a - 1, b) - MD(eL, a - 1, b)) * coeff_i;
        return minmod2d(central_u * coeff_u, d_r_y, d_l_y, d_r_x, d_l_x) / coeff_u;
      };
      int done = 0;
      for (int a = M; a >= 2; --a) {
        double limited = limiting(a, a);
        if (limited != MD(e, a, a)) UN(a, a) = limited;
        else break;
        for (int b = a - 1; b >= 2; --b) {
          double l1 = limiting(a, b), l2 = limiting(b, a);
          if ((fabs(l1 - MD(e, a, b)) < P.eps) && (fabs(l2 - MD(e, b, a)) < P.eps)) { done = 1; break; }
          UN(a, b) = l1;
          UN(b, a) = l2;
        }
        if (done == 1) break;
        double coeff_y = (2 * (double)(a - 1) + 1), coeff_u = (2 * (double)(a - 1) + 1);
        double d_r_y = MD(eT, a - 1, 1) - MD(e, a - 1, 1);
        double d_l_y = MD(e, a - 1, 1) - MD(eB, a - 1, 1);
        double d_r_x = MD(eR, 1, a - 1) - MD(e, 1, a - 1);
        double d_l_x = MD(e, 1, a - 1) - MD(eL, 1, a - 1);
        double l1 = generalized_minmod(P, MD(e, 1, a) * coeff_u, d_r_y * coeff_y, d_l_y * coeff_y) / coeff_u;
        double l2 = generalized_minmod(P, MD(e, a, 1) * coeff_u, d_r_x * coeff_y, d_l_x * coeff_y) / coeff_u;
        if ((l1 == MD(e, 1, a)) && (l2 == MD(e, a, 1))) break;
        UN(1, a) = l1;
        UN(a, 1) = l2;
      }
    }
#undef MD
#undef UN
    // compute_positivity(u_new) :1571-1574.  Sufficient test first (as in the fused stage kernels): every point value of a
    // variable lies within R_v = sum |mode_ij| P_i(1) P_j(1) of its mean; bounds of density and pressure above eps with a
    // margin far above the rounding of the point evaluations mean theta = 1 and t = 1 exactly, i.e. nothing to do
    bool ok = false;
    {
      double R[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        double r = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = 0; j < M; ++j)
            if (i != 0 || j != 0) r = r + fabs(un[v][i][j]) * (B.Ep[i] * B.Ep[j]);
        R[v] = r;
      }
      const double rho_lo = un[0][0][0] - R[0], E_lo = un[3][0][0] - R[3];
      const double mx_hi = fabs(un[1][0][0]) + R[1], my_hi = fabs(un[2][0][0]) + R[2];
      const double margin = 1e-9 * (fabs(un[0][0][0]) + R[0] + fabs(un[3][0][0]) + R[3]);
      if (rho_lo > P.eps + margin && rho_lo > (double)10e-10f) {
        const double p_lo = P.gm1a * (E_lo - 0.5 * (mx_hi * mx_hi + my_hi * my_hi) / rho_lo);
        ok = p_lo > P.eps + margin;
      }
    }
    if (!ok) positivity_el<M>(P, B, un);
  }
  store_modes<M>(out, g, e, un);
  if (out2) {
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int i = 0; i < M; ++i) PL(out2, g, v, j * M + i)[e] = fma(k3, un[v][i][j], PL(out2, g, v, j * M + i)[e]);
  }
}
// limiter_low_order ('LOW') :769-860 as a copy u -> out
template <int M>
__global__ void k_limiter_low_into(const double* __restrict__ u, double* __restrict__ out, DgGrid g, const DgCtrl* __restrict__ ctrl) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne) return;
  const bool act = !ctrl->skip && M > 1;
  for (int v = 0; v < 4; ++v)
    for (int j = 0; j < M; ++j)
      for (int i = 0; i < M; ++i) {
        const bool zero = act && ((i == 0 && j >= 1) || (j == 0 && i >= 1));
        PL(out, g, v, j * M + i)[e] = zero ? 0.0 : PL(u, g, v, j * M + i)[e];
      }
}
// out2 += k3 * out (the k3*w4 term of w5), and the pass-through of a skipped step for limiters that write `out` themselves
__global__ void k_dg_finish_out2(double* __restrict__ out2, double k3, const double* __restrict__ lim, size_t n,
                                 const DgCtrl* __restrict__ ctrl) {
  if (ctrl->skip) return;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    out2[k] = fma(k3, lim[k], out2[k]);
}

// compute_limiter ('1OR') :203-309, step A: modal PRIMITIVE variables w = modes(prim(nodes(u)))
template <int M>
__global__ void k_limiter_1or_a(const double* __restrict__ u, double* __restrict__ w, DgGrid g, DgPhys P, Basis B,
                                const DgCtrl* __restrict__ ctrl) {
  if (ctrl && ctrl->skip) return;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne) return;
  double md[4][M][M], nd[4][M][M];
  load_modes<M>(u, g, e, md);
  nodes_from_modes_el<M>(md, B, nd);
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double uu[4] = {nd[0][i][j], nd[1][i][j], nd[2][i][j], nd[3][i][j]}, ww[4];
      prim(P, uu, ww);
#pragma unroll
      for (int v = 0; v < 4; ++v) nd[v][i][j] = ww[v];
    }
  modes_from_nodes_el<M>(nd, B, md);
  store_modes<M>(w, g, e, md);
}
// step B: minmod on the linear modes against the neighbours' means, drop the higher modes, back to conservative modes
template <int M>
__global__ void k_limiter_1or_b(const double* __restrict__ w, double* __restrict__ u, DgGrid g, DgPhys P, Basis B,
                                const DgCtrl* __restrict__ ctrl) {
  if (ctrl && ctrl->skip) return;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne) return;
  const int ic = (int)(e % g.nx), jc = (int)(e / g.nx);
  const size_t eL = (size_t)jc * g.nx + bc_index(P.bc, ic - 1, g.nx), eR = (size_t)jc * g.nx + bc_index(P.bc, ic + 1, g.nx);
  const size_t eB = (size_t)y_nb(g, P.bc, jc - 1) * g.nx + ic, eT = (size_t)y_nb(g, P.bc, jc + 1) * g.nx + ic;
  const double norm = 3.;
  double md[4][M][M], nd[4][M][M];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) md[v][i][j] = 0.0;
    const double c = PL(w, g, v, 0)[e];
    md[v][0][0] = c;
    if (M > 1) {
      const double u21 = PL(w, g, v, 1)[e], u12 = PL(w, g, v, M)[e];
      double l1 = generalized_minmod(P, norm * u21, (PL(w, g, v, 0)[eR] - c), (c - PL(w, g, v, 0)[eL])) / norm;
      double l2 = generalized_minmod(P, norm * u12, (PL(w, g, v, 0)[eT] - c), (c - PL(w, g, v, 0)[eB])) / norm;
      if ((fabs(l1 - u21) > (double)1E-6f) || fabs(l2 - u12) > (double)1E-6f) { md[v][1][0] = l1; md[v][0][1] = l2; }
      else { md[v][0][1] = u12; md[v][1][0] = u21; }
    }
  }
  nodes_from_modes_el<M>(md, B, nd);
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double ww[4] = {nd[0][i][j], nd[1][i][j], nd[2][i][j], nd[3][i][j]}, uu[4];
      cons(P, ww, uu);
#pragma unroll
      for (int v = 0; v < 4; ++v) nd[v][i][j] = uu[v];
    }
  modes_from_nodes_el<M>(nd, B, md);
  store_modes<M>(u, g, e, md);
}

// limiter_positivity ('POS', 2d/limiters.f90:863-1036), reference operation order, one thread per element, u -> out.
// minmod on the two linear modes of the conserved variables against the CLAMPED neighbour means of the unlimited input
// (the clamp ignores bc; the reference writes the y pass with the cell indices swapped and runs both loops to nx), then
// the nodal primitive density / pressure are reset to the real(4) literal 1e-5 wherever the first or last node of their
// row or column is below 1d-10 -- the loop re-reads nodes it may already have reset, so it is kept sequential -- and the
// element goes back through compute_conservative and the projection.
template <int M>
__global__ void k_limiter_pos(const double* __restrict__ u, double* __restrict__ out, DgGrid g, DgPhys P, Basis B,
                              const DgCtrl* __restrict__ ctrl) {
  if (ctrl && ctrl->skip) return;
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ne) return;
  const int ic = (int)(e % g.nx), jc = (int)(e / g.nx);
  const size_t eL = (size_t)jc * g.nx + max(ic - 1, 0), eR = (size_t)jc * g.nx + min(ic + 1, g.nx - 1);
  // the clamp is on the GLOBAL row (it ignores bc): on a slab the ghost rows hold the neighbours' rows, except at the ends
  // of a periodic box, where the reference takes the element itself
  const int gj = g.j0 + jc;
  const size_t eB = (size_t)(gj - 1 < 0 ? jc : max(jc - 1, 0)) * g.nx + ic;
  const size_t eT = (size_t)(gj + 1 > g.nyg - 1 ? jc : min(jc + 1, g.ny - 1)) * g.nx + ic;
  double md[4][M][M], nd[4][M][M];
  load_modes<M>(u, g, e, md);
  if (M > 1) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const double u_center = 0.5 * md[v][0][0];
      {   // x pass (:899-935): mode (2,1), drops u(2:mx,1)
        const double u_left = 0.5 * PL(u, g, v, 0)[eL], u_right = 0.5 * PL(u, g, v, 0)[eR], u_deriv = md[v][1][0];
        const double l = minmod(u_deriv, (u_center - u_left) / P.dx, (u_right - u_center) / P.dx);
        const bool drop = fabs(l - u_deriv) > (double)0.01f * fabs(u_deriv);
        md[v][1][0] = l;
        if (drop) {
#pragma unroll
          for (int i = 1; i < M; ++i) md[v][i][0] = 0.0;
        }
      }
      {   // y pass (:941-978): mode (1,2), drops u(1,2:my)
        const double u_left = 0.5 * PL(u, g, v, 0)[eB], u_right = 0.5 * PL(u, g, v, 0)[eT], u_deriv = md[v][0][1];
        const double l = minmod(u_deriv, (u_center - u_left) / P.dx, (u_right - u_center) / P.dx);
        const bool drop = fabs(l - u_deriv) > (double)0.01f * fabs(u_deriv);
        md[v][0][1] = l;
        if (drop) {
#pragma unroll
          for (int j = 1; j < M; ++j) md[v][0][j] = 0.0;
        }
      }
    }
  }
  nodes_from_modes_el<M>(md, B, nd);
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double uu[4] = {nd[0][i][j], nd[1][i][j], nd[2][i][j], nd[3][i][j]}, ww[4];
      prim(P, uu, ww);
#pragma unroll
      for (int v = 0; v < 4; ++v) nd[v][i][j] = ww[v];
    }
#pragma unroll
  for (int v = 0; v < 4; v += 3)      // density and pressure only (:996-1012)
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j) {
        const double u_left = nd[v][0][j], u_right = nd[v][M - 1][j], u_top = nd[v][i][0], u_bottom = nd[v][i][M - 1];
        if (u_left < 1e-10 || u_right < 1e-10 || u_top < 1e-10 || u_bottom < 1e-10) nd[v][i][j] = (double)1e-5f;
      }
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double ww[4] = {nd[0][i][j], nd[1][i][j], nd[2][i][j], nd[3][i][j]}, uu[4];
      cons(P, ww, uu);
#pragma unroll
      for (int v = 0; v < 4; ++v) nd[v][i][j] = uu[v];
    }
  modes_from_nodes_el<M>(nd, B, md);
  store_modes<M>(out, g, e, md);
}

// compute_error :23-89.  Per element the reference's accumulators (same order, both directions weighted with w_x_quad,
// :52) -- the per-variable sums over the elements are order dependent in the reference (i outer, j inner); here they are
// a fixed-shape tree (grid-stride partial sums per thread, block tree, one final block), i.e. deterministic and equal to
// the reference to a few ulp of the sum.  part: [gridDim.x][12] = lmax[4], l1[4], l2[4].
template <int M>
__global__ void k_dg_error(const double* __restrict__ u, const double* __restrict__ u0, DgGrid g, Basis B, double scale,
                           double* __restrict__ part) {
  double acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.0;
  for (size_t e = g.e_off + (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < g.e_off + g.ne_own; e += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      double a1 = 0.0, a2 = 0.0, m = acc[v];
#pragma unroll
      for (int qi = 0; qi < M; ++qi)
#pragma unroll
        for (int qj = 0; qj < M; ++qj) {
          const double d = PL(u, g, v, qj * M + qi)[e] - PL(u0, g, v, qj * M + qi)[e];
          a1 = a1 + fabs(d) * B.wq[qi] * B.wq[qj];
          a2 = a2 + d * d * B.wq[qi] * B.wq[qj];
          m = fmax(m, fabs(d));
        }
      acc[v] = m;
      acc[4 + v] = acc[4 + v] + a1 * scale * 0.25;
      acc[8 + v] = acc[8 + v] + a2 * scale * 0.25;
    }
  }
  __shared__ double sh[128][12];
#pragma unroll
  for (int k = 0; k < 12; ++k) sh[threadIdx.x][k] = acc[k];
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 12; ++k)
        sh[threadIdx.x][k] = (k < 4) ? fmax(sh[threadIdx.x][k], sh[threadIdx.x + s][k]) : sh[threadIdx.x][k] + sh[threadIdx.x + s][k];
    __syncthreads();
  }
  if (threadIdx.x < 12) part[(size_t)blockIdx.x * 12 + threadIdx.x] = sh[0][threadIdx.x];
}
__global__ void k_dg_error_final(const double* __restrict__ part, int nparts, double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= 12) return;
  double r = 0.0;
  for (int b = 0; b < nparts; ++b) r = (k < 4) ? fmax(r, part[(size_t)b * 12 + k]) : r + part[(size_t)b * 12 + k];
  out[k] = r;
}

// ------------------------------------------------------------------------------------ compute_max_speed :826-870
// The reference scan (i outer, j inner; `>=` keeps the LAST maximum; every later cell with a smaller cs lowers
// cs_max) in its commutative two-phase form:
//   phase 1: (speed_max, k*) = lexicographic max of (speed, k), k = i*ny + j; v_x, v_y, cs taken at k*
//   phase 2: cs_max = min(cs(k*), min of cs over k > k*)
struct SpeedKey { double speed; long long k; double vx, vy, cs; };
__device__ __forceinline__ bool key_less(const SpeedKey& a, const SpeedKey& b) {   // a < b
  return (a.speed < b.speed) || (a.speed == b.speed && a.k < b.k);
}
__global__ void k_speed_phase1(const double* __restrict__ u, DgGrid g, DgPhys P, SpeedKey* __restrict__ part) {
  SpeedKey best{-1.0, -1, 0.0, 0.0, 0.0};
  for (size_t e = g.e_off + (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < g.e_off + g.ne_own; e += (size_t)gridDim.x * blockDim.x) {
    const int ic = (int)(e % g.nx), jc = (int)(e / g.nx);
    double uu[4] = {PL(u, g, 0, 0)[e], PL(u, g, 1, 0)[e], PL(u, g, 2, 0)[e], PL(u, g, 3, 0)[e]};
    SpeedKey c;
    speed(P, uu, c.cs, c.vx, c.vy, c.speed);
    c.k = (long long)ic * g.nyg + (g.j0 + jc);          // position in the reference's scan (i outer, j inner), global
    if (key_less(best, c)) best = c;
  }
  __shared__ SpeedKey sh[256];
  sh[threadIdx.x] = best;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s && key_less(sh[threadIdx.x], sh[threadIdx.x + s])) sh[threadIdx.x] = sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void k_speed_phase1b(const SpeedKey* __restrict__ part, int nparts, DgCtrl* ctrl) {
  __shared__ SpeedKey sh[256];
  SpeedKey best{-1.0, -1, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < nparts; i += blockDim.x)
    if (key_less(best, part[i])) best = part[i];
  sh[threadIdx.x] = best;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s && key_less(sh[threadIdx.x], sh[threadIdx.x + s])) sh[threadIdx.x] = sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // speed_max starts at 0.0 and the test is `speed >= speed_max`: with non-negative speeds the first cell always enters
    ctrl->speed_max = fmax(0.0, sh[0].speed);
    ctrl->vx = sh[0].vx; ctrl->vy = sh[0].vy; ctrl->cs_max = sh[0].cs; ctrl->kstar = sh[0].k;
  }
}
__global__ void k_speed_phase2(const double* __restrict__ u, DgGrid g, DgPhys P, const DgCtrl* __restrict__ ctrl,
                               double* __restrict__ part) {
  const long long kstar = ctrl->kstar;
  double m = ctrl->cs_max;
  for (size_t e = g.e_off + (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < g.e_off + g.ne_own; e += (size_t)gridDim.x * blockDim.x) {
    const int ic = (int)(e % g.nx), jc = (int)(e / g.nx);
    if ((long long)ic * g.nyg + (g.j0 + jc) <= kstar) continue;
    double uu[4] = {PL(u, g, 0, 0)[e], PL(u, g, 1, 0)[e], PL(u, g, 2, 0)[e], PL(u, g, 3, 0)[e]};
    double cs, vx, vy, sp;
    speed(P, uu, cs, vx, vy, sp);
    m = fmin(m, cs);
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = m;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = fmin(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
// final: cs_max and the time step of evolve (:671) -- dt = min(tend-t, cfl*min(1/9, gll_w_1/2)/((|vx|+cs)/dx + (|vy|+cs)/dx))
__device__ __forceinline__ void speed_finish(DgCtrl* ctrl, const DgPhys& P, int set_dt) {
  if (set_dt) {
    const bool done = !(ctrl->t < ctrl->tend) || (ctrl->max_iter >= 0 && ctrl->iter >= ctrl->max_iter);
    ctrl->skip = done ? 1 : 0;
    if (!done) {
      const double cs = ctrl->cs_max;
      ctrl->dt = fmin(ctrl->tend - ctrl->t, P.dt_num / ((fabs(ctrl->vx) + (cs)) / P.dx + (fabs(ctrl->vy) + (cs)) / P.dx));
    }
  }
}
__global__ void k_speed_phase2b(const double* __restrict__ part, int nparts, DgCtrl* ctrl, DgPhys P, int set_dt) {
  __shared__ double sh[256];
  double m = ctrl->cs_max;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) m = fmin(m, part[i]);
  sh[threadIdx.x] = m;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = fmin(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ctrl->cs_max = sh[0];
    if (set_dt >= 0) speed_finish(ctrl, P, set_dt);      // slab mode: set_dt < 0, the minimum is all-reduced first
  }
}
__global__ void k_speed_finish(DgCtrl* ctrl, DgPhys P, int set_dt) { speed_finish(ctrl, P, set_dt); }

// ---- slab mode: the two phases of the scan across ranks (lexicographic max of (speed, k), then the
//      minimum of cs over k >= k*), each a tiny all-reduce on `red`
__global__ void k_speed_red_a(const DgCtrl* ctrl, double* red) { red[0] = ctrl->speed_max; }
__global__ void k_speed_red_b(const DgCtrl* ctrl, double* red) {      // red[0] = global max speed -> my candidate k
  red[1] = (ctrl->speed_max == red[0]) ? (double)ctrl->kstar : -1.0;
}
__global__ void k_speed_red_c(const DgCtrl* ctrl, double* red) {      // red[1] = global k* -> the owner publishes its state
  const bool owner = (ctrl->speed_max == red[0]) && ((double)ctrl->kstar == red[1]);
  red[2] = owner ? ctrl->vx : -1.7976931348623157e308;
  red[3] = owner ? ctrl->vy : -1.7976931348623157e308;
  red[4] = owner ? ctrl->cs_max : -1.7976931348623157e308;
}
__global__ void k_speed_red_d(DgCtrl* ctrl, const double* red) {
  ctrl->speed_max = red[0]; ctrl->kstar = (long long)red[1]; ctrl->vx = red[2]; ctrl->vy = red[3]; ctrl->cs_max = red[4];
}

// ---- slab mode: ghost rows.  pack: first and last owned row of all 4*nm planes -> contiguous send buffers;
//      unpack: receive buffers -> ghost rows (row 0 and row ny-1)
__global__ void k_dg_pack_rows(const double* __restrict__ u, DgGrid g, double* __restrict__ lo, double* __restrict__ hi) {
  const int ic = blockIdx.x * blockDim.x + threadIdx.x, pl = blockIdx.y;
  if (ic >= g.nx) return;
  const double* p = u + (size_t)pl * g.ne;
  lo[(size_t)pl * g.nx + ic] = p[(size_t)1 * g.nx + ic];
  hi[(size_t)pl * g.nx + ic] = p[(size_t)(g.ny - 2) * g.nx + ic];
}
__global__ void k_dg_unpack_rows(double* __restrict__ u, DgGrid g, const double* __restrict__ lo, const double* __restrict__ hi) {
  const int ic = blockIdx.x * blockDim.x + threadIdx.x, pl = blockIdx.y;
  if (ic >= g.nx) return;
  double* p = u + (size_t)pl * g.ne;
  p[ic] = lo[(size_t)pl * g.nx + ic];
  p[(size_t)(g.ny - 1) * g.nx + ic] = hi[(size_t)pl * g.nx + ic];
}
__global__ void k_dg_copy_if_skipped(double* __restrict__ out, const double* __restrict__ in, size_t n, const DgCtrl* __restrict__ ctrl) {
  if (!ctrl->skip) return;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) out[k] = in[k];
}
__global__ void k_dg_advance(DgCtrl* ctrl) {
  if (ctrl->skip) return;
  ctrl->t = ctrl->t + ctrl->dt;
  ctrl->iter = ctrl->iter + 1;
}
__global__ void k_dg_ctrl_init(DgCtrl* ctrl, double tend, int max_iter, int reset_clock) {
  if (reset_clock) { ctrl->t = 0.0; ctrl->iter = 0; ctrl->dt = 0.0; }
  ctrl->tend = tend; ctrl->max_iter = max_iter; ctrl->skip = 0;
}

}}  // namespace wb::dg

#include "dg2d_fast.cuh"
#include "dg2d_tma.cuh"

namespace wb { namespace dg {
// k_dg_stage_split lives in its own translation unit (dg2d_split.cu)
int launch_stage_split(const CUtensorMap* map, const double* in, const StageCoef& C, double* out, const double* gx,
                       const double* gy, const unsigned char* fz, const DgGrid& g, const DgPhys& P, const FastBasis& B,
                       const DgCtrl* ctrl, int onp, int rows, int row_begin, int row_end, cudaStream_t stream);
}}

// ============================================================================================ host side
using namespace wb;
using namespace wb::dg;

struct wb_dg2d {
  wb_dg2d_params prm;
  int dev = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  DgGrid g;
  DgPhys phys;
  Basis B;
  size_t nfield = 0;                 // doubles per field (4*nm*ne)
  double *du = nullptr, *A = nullptr, *Bf = nullptr, *C = nullptr, *D = nullptr, *E = nullptr;   // state, RK buffers, dudt, scratch
  double *gx = nullptr, *gy = nullptr, *stage = nullptr, *xy = nullptr;
  unsigned char* fz = nullptr;
  DgCtrl* ctrl = nullptr;
  DgCtrl* h_ctrl = nullptr;
  SpeedKey* part1 = nullptr;
  double* part2 = nullptr;
  int nparts = 0;
  bool resident = false;
  bool have_xy = false;
  FastBasis FB;
  int arith = 0;               // 0 = fused/sum-factorised stage kernel, 1 = reference operation order
  // slab mode (nranks > 1): NCCL communicator, packed boundary rows (send lo/hi, receive lo/hi), all-reduce scratch
  wb::Nccl* comm = nullptr;
  double *sbuf_lo = nullptr, *sbuf_hi = nullptr, *rbuf_lo = nullptr, *rbuf_hi = nullptr, *red = nullptr;
  int rank = 0, nranks = 1, nyl = 0;
  // overlap of the ghost exchange with the stage kernel: the two boundary rows, their pack / NCCL / unpack on a high-priority
  // stream while the interior rows run on the main one (WB_DG2D_OVERLAP=0 switches it off)
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_main = nullptr, ev_comm = nullptr;
  bool overlap = true;
  // ... and with the neighbours' four state buffers mapped (CUDA IPC) the boundary-row launches store their rows straight
  // into the neighbours' ghost rows: no pack / NCCL / unpack, one flag word per direction (common.cuh: peer_signal / peer_wait)
  struct Peer {
    bool active = false;
    double* lo[4] = {nullptr, nullptr, nullptr, nullptr};      // the lower neighbour's buffers, in the order of map_ptr[]
    double* hi[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long *lo_flags = nullptr, *hi_flags = nullptr, *flags = nullptr;
    size_t lo_ne = 0, hi_ne = 0;
    int lo_ny = 0, hi_ny = 0;
    bool same = false;                                         // two ranks on a ring: both neighbours are the one other rank
    unsigned long long seq = 0;
  } peer;
  // TMA-staged stage kernel: one 3-D tensor map (column, row, plane) per state buffer
  bool tma_ok = false;
  int march_rows = 32;         // rows per strip of k_dg_stage_split
  bool split_ok = false;       // k_dg_stage_split (element split over four threads, every face once): nx % 32 == 0
  const double* map_ptr[4] = {nullptr, nullptr, nullptr, nullptr};
  CUtensorMap map[4];
  wb::OutputJob* out_job = nullptr;   // output_file in flight (host thread)
};

namespace {

#define DISPATCH_M(h, ...)                              \
  switch ((h)->g.m) {                                   \
    case 1: { constexpr int MM = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int MM = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int MM = 3; __VA_ARGS__; } break; \
    default: { constexpr int MM = 4; __VA_ARGS__; } break; \
  }

inline dim3 elem_grid(const wb_dg2d* h, int block) { return dim3((unsigned)((h->g.ne + block - 1) / block)); }

int dg_ensure(wb_dg2d* h, double** buf) {
  if (!*buf) WB_CUDA(cudaMalloc(buf, sizeof(double) * h->nfield));
  return WB_OK;
}
// after k_grad_phi: decide once whether the stage kernel may read the gravity field from its separable lines
int dg_grad_sep(wb_dg2d* h) {
  h->phys.gsep = 0;
  h->phys.grow = h->g.slab ? 1 : 0;
  if (h->phys.source != 2) return WB_OK;
  const int row0 = h->phys.grow, row1 = h->g.slab ? h->g.ny - 1 : h->g.ny;
  int one = 1, *flag = reinterpret_cast<int*>(h->part2);      // scratch of the max-speed scan: free outside a step
  WB_CUDA(cudaMemcpyAsync(flag, &one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  dim3 b(128), gr((h->g.nx + 127) / 128, row1 - row0);
  k_grad_sep_check<<<gr, b, 0, h->stream>>>(h->gx, h->gy, h->g, row0, row1, flag);
  WB_LAUNCH_CHECK();
  WB_CUDA(cudaMemcpyAsync(&one, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (const char* e = getenv("WB_DG2D_GSEP")) one = one && atoi(e) != 0;
  h->phys.gsep = one;
  return WB_OK;
}
int dg_h2d_field(wb_dg2d* h, const double* host, double* soa) {
  WB_CHECK(dg_ensure(h, &h->stage));
  WB_CUDA(cudaMemcpyAsync(h->stage, host, sizeof(double) * 4 * h->g.nm * h->g.ne_own, cudaMemcpyHostToDevice, h->stream));
  dim3 b(128), gr((unsigned)((h->g.ne_own + 127) / 128), h->g.nm);
  k_dg_aos_to_soa<<<gr, b, 0, h->stream>>>(h->stage, soa, h->g);
  WB_LAUNCH_CHECK();
  return WB_OK;
}
int dg_d2h_field(wb_dg2d* h, const double* soa, double* host) {
  WB_CHECK(dg_ensure(h, &h->stage));
  dim3 b(128), gr((unsigned)((h->g.ne_own + 127) / 128), h->g.nm);
  k_dg_soa_to_aos<<<gr, b, 0, h->stream>>>(soa, h->stage, h->g);
  WB_LAUNCH_CHECK();
  WB_CUDA(cudaMemcpyAsync(host, h->stage, sizeof(double) * 4 * h->g.nm * h->g.ne_own, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  return WB_OK;
}
// x,y(nx,ny,mx,my) have the device layout already ([mode][jc][ic]); derive grad_phi / freeze mask from them
int dg_set_xy(wb_dg2d* h, const double* x, const double* y) {
  const size_t n = (size_t)h->g.nm * h->g.ne;
  const bool need = (h->phys.source == 2) || (h->phys.ninit == 12);
  h->have_xy = true;
  if (!need) return WB_OK;
  WB_REQUIRE(x && y, "x and y are required when source == 2 or ninit == 12");
  if (h->g.slab) WB_CUDA(cudaMemsetAsync(h->xy, 0, sizeof(double) * 2 * n, h->stream));      // ghost rows: finite coordinates
  for (int m = 0; m < h->g.nm; ++m) {         // host x,y hold the owned rows of every node plane
    WB_CUDA(cudaMemcpyAsync(h->xy + (size_t)m * h->g.ne + h->g.e_off, x + (size_t)m * h->g.ne_own, sizeof(double) * h->g.ne_own,
                            cudaMemcpyHostToDevice, h->stream));
    WB_CUDA(cudaMemcpyAsync(h->xy + n + (size_t)m * h->g.ne + h->g.e_off, y + (size_t)m * h->g.ne_own, sizeof(double) * h->g.ne_own,
                            cudaMemcpyHostToDevice, h->stream));
  }
  dim3 b(256), gr((unsigned)((n + 255) / 256));
  if (h->phys.source == 2) {
    k_grad_phi<<<gr, b, 0, h->stream>>>(h->xy, h->xy + n, h->gx, h->gy, n, h->prm.grad_phi_case);
    WB_LAUNCH_CHECK();
    WB_CHECK(dg_grad_sep(h));
  }
  if (h->phys.ninit == 12) {
    k_freeze_mask<<<gr, b, 0, h->stream>>>(h->xy, h->xy + n, h->fz, n, h->prm.boxlen_x / 2., h->prm.boxlen_y / 2.);
    WB_LAUNCH_CHECK();
  }
  return WB_OK;
}

int dg_update(wb_dg2d* h, const double* in, double* out, bool use_ctrl) {
  dim3 b(128), gr = elem_grid(h, 128);
  DISPATCH_M(h, k_dg_update<MM><<<gr, b, 0, h->stream>>>(in, h->gx, h->gy, h->phys.ninit == 12 ? h->fz : nullptr, out, h->g,
                                                        h->phys, h->B, use_ctrl ? h->ctrl : nullptr));
  WB_LAUNCH_CHECK();
  return WB_OK;
}

int dg_limiter(wb_dg2d* h, double* u, bool use_ctrl) {
  const DgCtrl* c = use_ctrl ? h->ctrl : nullptr;
  if (h->prm.limiter_id == 2 || h->prm.limiter_id == 3 || h->prm.limiter_id >= 5) WB_CHECK(dg_ensure(h, &h->E));
  dim3 b(128), gr = elem_grid(h, 128);
  if (h->g.m == 1) return WB_OK;           // every limiter returns early for mx == my == 1
  switch (h->prm.limiter_id) {
    case 1:
      DISPATCH_M(h, k_limiter_onp<MM><<<gr, b, 0, h->stream>>>(u, h->g, h->phys, h->B, c));
      WB_LAUNCH_CHECK();
      break;
    case 2:
      DISPATCH_M(h, k_limiter_hio<MM><<<gr, b, 0, h->stream>>>(u, h->E, h->g, h->phys, c));
      WB_LAUNCH_CHECK();
      // u = u_new, then compute_positivity(u) (:1571-1574); a skipped step must not copy a stale scratch buffer
      DISPATCH_M(h, k_limiter_onp<MM><<<gr, b, 0, h->stream>>>(h->E, h->g, h->phys, h->B, c));
      WB_LAUNCH_CHECK();
      {
        dim3 gb((unsigned)std::min<size_t>((h->nfield + 255) / 256, 148 * 16));
        k_dg_axpy<0><<<gb, 256, 0, h->stream>>>(u, h->E, 1.0, nullptr, 0, nullptr, 0, nullptr, 0, h->E, 0.0, h->nfield,
                                                c ? c : h->ctrl);
        WB_LAUNCH_CHECK();
      }
      break;
    case 3:
      DISPATCH_M(h, k_limiter_1or_a<MM><<<gr, b, 0, h->stream>>>(u, h->E, h->g, h->phys, h->B, c));
      WB_LAUNCH_CHECK();
      DISPATCH_M(h, k_limiter_1or_b<MM><<<gr, b, 0, h->stream>>>(h->E, u, h->g, h->phys, h->B, c));
      WB_LAUNCH_CHECK();
      break;
    case 4:
      DISPATCH_M(h, k_limiter_low<MM><<<gr, b, 0, h->stream>>>(u, h->g, c));
      WB_LAUNCH_CHECK();
      break;
    case 5:
      DISPATCH_M(h, k_limiter_pos<MM><<<gr, b, 0, h->stream>>>(u, h->E, h->g, h->phys, h->B, c));
      WB_LAUNCH_CHECK();
      {     // u = u_lim (:1035); a skipped step must not copy a stale scratch buffer
        dim3 gb((unsigned)std::min<size_t>((h->nfield + 255) / 256, 148 * 16));
        k_dg_axpy<0><<<gb, 256, 0, h->stream>>>(u, h->E, 1.0, nullptr, 0, nullptr, 0, nullptr, 0, h->E, 0.0, h->nfield,
                                                c ? c : h->ctrl);
        WB_LAUNCH_CHECK();
      }
      break;
    case 6:      // pass B reads only the element's own modes of u (the neighbours through the scratch field): in place
      DISPATCH_M(h, k_limiter_po3_a<MM><<<gr, b, 0, h->stream>>>(u, h->E, h->g, h->phys, h->B, c));
      WB_LAUNCH_CHECK();
      DISPATCH_M(h, k_limiter_po3_b<MM><<<gr, b, 0, h->stream>>>(u, h->E, u, h->g, h->phys, h->B, c));
      WB_LAUNCH_CHECK();
      break;
    default: break;
  }
  return WB_OK;
}

// fused flow with a neighbour-reading limiter: `tmp` (un-limited stage result, ghost rows valid) -> `out`; out2 gets its
// k3*out term.  1OR and POS keep their own kernels (1OR needs the primitive modes of the neighbours: two passes).
int dg_limit_into(wb_dg2d* h, double* tmp, double* out, double* out2, double k3) {
  dim3 b(128), gr = elem_grid(h, 128);
  dim3 gb((unsigned)std::min<size_t>((h->nfield + 255) / 256, 148 * 16));
  if (h->g.m == 1) {                          // every limiter returns early for mx == my == 1: out = tmp
    k_dg_axpy<0><<<gb, 256, 0, h->stream>>>(out, tmp, 1.0, nullptr, 0, nullptr, 0, nullptr, 0, tmp, 0.0, h->nfield, h->ctrl);
    k_dg_copy_if_skipped<<<gb, 256, 0, h->stream>>>(out, tmp, h->nfield, h->ctrl);
    wb::g_launches.fetch_add(1);
    WB_LAUNCH_CHECK();
  } else switch (h->prm.limiter_id) {
    case 2:
      DISPATCH_M(h, k_limiter_hio_onp<MM><<<gr, b, 0, h->stream>>>(tmp, out, out2, k3, h->g, h->phys, h->B, h->ctrl));
      WB_LAUNCH_CHECK();
      return WB_OK;
    case 3:      // the ghost rows of E come from the (exchanged) ghost rows of tmp: step A is element-local
      WB_CHECK(dg_ensure(h, &h->E));
      DISPATCH_M(h, k_limiter_1or_a<MM><<<gr, b, 0, h->stream>>>(tmp, h->E, h->g, h->phys, h->B, h->ctrl));
      WB_LAUNCH_CHECK();
      k_dg_copy_if_skipped<<<gb, 256, 0, h->stream>>>(out, tmp, h->nfield, h->ctrl);      // the kernels return early on a skipped step
      WB_LAUNCH_CHECK();
      DISPATCH_M(h, k_limiter_1or_b<MM><<<gr, b, 0, h->stream>>>(h->E, out, h->g, h->phys, h->B, h->ctrl));
      WB_LAUNCH_CHECK();
      break;
    case 4:
      DISPATCH_M(h, k_limiter_low_into<MM><<<gr, b, 0, h->stream>>>(tmp, out, h->g, h->ctrl));
      WB_LAUNCH_CHECK();
      break;
    case 5:
      k_dg_copy_if_skipped<<<gb, 256, 0, h->stream>>>(out, tmp, h->nfield, h->ctrl);
      WB_LAUNCH_CHECK();
      DISPATCH_M(h, k_limiter_pos<MM><<<gr, b, 0, h->stream>>>(tmp, out, h->g, h->phys, h->B, h->ctrl));
      WB_LAUNCH_CHECK();
      break;
    case 6:
      WB_CHECK(dg_ensure(h, &h->E));
      DISPATCH_M(h, k_limiter_po3_a<MM><<<gr, b, 0, h->stream>>>(tmp, h->E, h->g, h->phys, h->B, h->ctrl));
      WB_LAUNCH_CHECK();
      k_dg_copy_if_skipped<<<gb, 256, 0, h->stream>>>(out, tmp, h->nfield, h->ctrl);
      WB_LAUNCH_CHECK();
      DISPATCH_M(h, k_limiter_po3_b<MM><<<gr, b, 0, h->stream>>>(tmp, h->E, out, h->g, h->phys, h->B, h->ctrl));
      WB_LAUNCH_CHECK();
      break;
    default:
      set_error("dg_limit_into: limiter %d has no fused flow", h->prm.limiter_id);
      return WB_ERR_STATE;
  }
  if (out2) {
    k_dg_finish_out2<<<gb, 256, 0, h->stream>>>(out2, k3, out, h->nfield, h->ctrl);
    WB_LAUNCH_CHECK();
  }
  return WB_OK;
}

// slab mode: fill the two ghost rows of a 4*nm-plane field.  Periodic box (bc = 1): ring of ranks; index clamp
// (bc = 2, 3): chain, and the ghost row at a global edge is the rank's own boundary row (the clamped neighbour).
int dg_exchange(wb_dg2d* h, double* field, cudaStream_t stream = nullptr) {
  if (!h->g.slab) return WB_OK;
  if (!stream) stream = h->stream;
  if (h->nranks > 1 && !h->comm) { set_error("nranks > 1 but wb_dg2d_comm_init was not called"); return WB_ERR_STATE; }
  const int npl = 4 * h->g.nm;
  const size_t cnt = (size_t)npl * h->g.nx;
  dim3 b(128), gr((h->g.nx + 127) / 128, npl);
  k_dg_pack_rows<<<gr, b, 0, stream>>>(field, h->g, h->sbuf_lo, h->sbuf_hi);
  WB_LAUNCH_CHECK();
  const bool periodic = (h->phys.bc == 1);
  const int lo_peer = (h->rank > 0) ? h->rank - 1 : (periodic ? h->nranks - 1 : -1);
  const int hi_peer = (h->rank < h->nranks - 1) ? h->rank + 1 : (periodic ? 0 : -1);
  WB_CHECK(nccl_ring_exchange(h->comm, lo_peer, hi_peer, h->sbuf_lo, h->sbuf_hi, h->rbuf_lo, h->rbuf_hi, cnt, stream));
  k_dg_unpack_rows<<<gr, b, 0, stream>>>(field, h->g, lo_peer >= 0 ? h->rbuf_lo : h->sbuf_lo, hi_peer >= 0 ? h->rbuf_hi : h->sbuf_hi);
  WB_LAUNCH_CHECK();
  return WB_OK;
}

// max speed of the mean mode of `u` into ctrl (and, if set_dt, the step's dt / skip flag)
int dg_max_speed(wb_dg2d* h, const double* u, int set_dt) {
  const bool multi = h->nranks > 1;
  k_speed_phase1<<<h->nparts, 256, 0, h->stream>>>(u, h->g, h->phys, h->part1);
  WB_LAUNCH_CHECK();
  k_speed_phase1b<<<1, 256, 0, h->stream>>>(h->part1, h->nparts, h->ctrl);
  WB_LAUNCH_CHECK();
  if (multi) {      // (speed, k) lexicographic max over ranks, then the owner's (vx, vy, cs)
    if (!h->comm) { set_error("nranks > 1 but wb_dg2d_comm_init was not called"); return WB_ERR_STATE; }
    k_speed_red_a<<<1, 1, 0, h->stream>>>(h->ctrl, h->red);
    WB_CHECK(nccl_allreduce_max_f64(h->comm, h->red, 1, h->stream));
    k_speed_red_b<<<1, 1, 0, h->stream>>>(h->ctrl, h->red);
    WB_CHECK(nccl_allreduce_max_f64(h->comm, h->red + 1, 1, h->stream));
    k_speed_red_c<<<1, 1, 0, h->stream>>>(h->ctrl, h->red);
    WB_CHECK(nccl_allreduce_max_f64(h->comm, h->red + 2, 3, h->stream));
    k_speed_red_d<<<1, 1, 0, h->stream>>>(h->ctrl, h->red);
    wb::g_launches.fetch_add(4);
    WB_CUDA(cudaGetLastError());
  }
  k_speed_phase2<<<h->nparts, 256, 0, h->stream>>>(u, h->g, h->phys, h->ctrl, h->part2);
  WB_LAUNCH_CHECK();
  k_speed_phase2b<<<1, 256, 0, h->stream>>>(h->part2, h->nparts, h->ctrl, h->phys, multi ? -1 : set_dt);
  WB_LAUNCH_CHECK();
  if (multi) {
    WB_CHECK(nccl_allreduce_min_f64(h->comm, &h->ctrl->cs_max, 1, h->stream));
    k_speed_finish<<<1, 1, 0, h->stream>>>(h->ctrl, h->phys, set_dt);
    WB_LAUNCH_CHECK();
  }
  return WB_OK;
}

// real(4) literal promoted to real(8)
#define F32(x) ((double)(x##f))

int dg_axpy(wb_dg2d* h, int na, double* out, const double* A0, double c0, const double* A1, double c1, const double* A2,
            double c2, const double* A3, double c3, const double* D, double cd) {
  dim3 gb((unsigned)std::min<size_t>((h->nfield + 255) / 256, 148 * 16));
  if (na == 1) k_dg_axpy<1><<<gb, 256, 0, h->stream>>>(out, A0, c0, A1, c1, A2, c2, A3, c3, D, cd, h->nfield, h->ctrl);
  else if (na == 2) k_dg_axpy<2><<<gb, 256, 0, h->stream>>>(out, A0, c0, A1, c1, A2, c2, A3, c3, D, cd, h->nfield, h->ctrl);
  else k_dg_axpy<4><<<gb, 256, 0, h->stream>>>(out, A0, c0, A1, c1, A2, c2, A3, c3, D, cd, h->nfield, h->ctrl);
  WB_LAUNCH_CHECK();
  return WB_OK;
}

// 3-D tensor map (column, local row, plane) of a state buffer for k_dg_stage_split
int dg_make_map(const wb_dg2d* h, const double* base, CUtensorMap* out) {
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
  const DgGrid& g = h->g;
  const cuuint64_t dims[3] = {(cuuint64_t)g.nx, (cuuint64_t)g.ny, (cuuint64_t)(4 * g.nm)};
  const cuuint64_t strides[2] = {(cuuint64_t)g.nx * sizeof(double), (cuuint64_t)g.ne * sizeof(double)};
  const cuuint32_t box[3] = {(cuuint32_t)DGT_W, 1u, (cuuint32_t)(4 * g.nm)};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return WB_ERR_CUDA; }
  return WB_OK;
}

// ---- peer-memory ghost rows (see wb_dg2d::Peer) -----------------------------------------------------------------
int dg_buffer_index(const wb_dg2d* h, const double* p) {
  for (int k = 0; k < 4; ++k)
    if (h->map_ptr[k] == p) return k;
  return -1;
}
// ghost-row targets of `out` (and `out2`) on the neighbours, or on the rank itself at a clamped global edge
bool dg_peer_targets(const wb_dg2d* h, const double* out, const double* out2, StageCoef& C) {
  const wb_dg2d::Peer& P = h->peer;
  const int k = dg_buffer_index(h, out), k2 = out2 ? dg_buffer_index(h, out2) : -1;
  if (k < 0 || (out2 && k2 < 0)) return false;
  const size_t nx = h->g.nx;
  // lower side: the neighbour's TOP ghost row (its local row ny-1); clamped edge: my own row 0
  double* lo = P.lo[k] ? P.lo[k] : const_cast<double*>(h->map_ptr[k]);
  double* lo2 = out2 ? (P.lo[k2] ? P.lo[k2] : const_cast<double*>(h->map_ptr[k2])) : nullptr;
  const size_t lo_row = P.lo[k] ? (size_t)(P.lo_ny - 1) : 0;
  C.peer_lo_ne = P.lo[k] ? P.lo_ne : h->g.ne;
  C.peer_lo = lo + lo_row * nx;
  C.peer2_lo = lo2 ? lo2 + lo_row * nx : nullptr;
  // upper side: the neighbour's BOTTOM ghost row (its local row 0); clamped edge: my own row ny-1
  double* hi = P.hi[k] ? P.hi[k] : const_cast<double*>(h->map_ptr[k]);
  double* hi2 = out2 ? (P.hi[k2] ? P.hi[k2] : const_cast<double*>(h->map_ptr[k2])) : nullptr;
  const size_t hi_row = P.hi[k] ? 0 : (size_t)(h->g.ny - 1);
  C.peer_hi_ne = P.hi[k] ? P.hi_ne : h->g.ne;
  C.peer_hi = hi + hi_row * nx;
  C.peer2_hi = hi2 ? hi2 + hi_row * nx : nullptr;
  return true;
}

struct DgPeerRecord {      // all-gathered once at comm_init
  cudaIpcMemHandle_t buf[4], flags;
  unsigned long long ne;
  int ny, ok;
};

void dg_peer_close(wb_dg2d* h) {
  wb_dg2d::Peer& P = h->peer;
  for (int k = 0; k < 4; ++k) {
    if (P.lo[k]) cudaIpcCloseMemHandle(P.lo[k]);
    if (P.hi[k] && !P.same) cudaIpcCloseMemHandle(P.hi[k]);
    P.lo[k] = P.hi[k] = nullptr;
  }
  if (P.lo_flags) cudaIpcCloseMemHandle(P.lo_flags);
  if (P.hi_flags && !P.same) cudaIpcCloseMemHandle(P.hi_flags);
  P.lo_flags = P.hi_flags = nullptr;
  P.active = false;
  cudaGetLastError();
}

// Collective over the communicator; any failure anywhere leaves every rank on pack / NCCL / unpack (all-reduced decision).
int dg_peer_setup(wb_dg2d* h) {
  wb_dg2d::Peer& P = h->peer;
  const int R = h->nranks, r = h->rank;
  const bool periodic = (h->phys.bc == 1);
  const int lo_rank = (r > 0) ? r - 1 : (periodic ? R - 1 : -1), hi_rank = (r < R - 1) ? r + 1 : (periodic ? 0 : -1);
  const char* env = getenv("WB_DG2D_P2P");
  int ok = (!env || atoi(env) != 0) && h->tma_ok && h->split_ok && h->overlap && h->comm_stream && h->g.ny >= 6 && h->arith == 0;
  DgPeerRecord mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && !P.flags) ok = cudaMalloc(&P.flags, 4 * sizeof(unsigned long long)) == cudaSuccess;
  if (ok) ok = cudaMemsetAsync(P.flags, 0, 4 * sizeof(unsigned long long), h->stream) == cudaSuccess;
  for (int k = 0; k < 4 && ok; ++k) ok = cudaIpcGetMemHandle(&mine.buf[k], const_cast<double*>(h->map_ptr[k])) == cudaSuccess;
  if (ok) ok = cudaIpcGetMemHandle(&mine.flags, P.flags) == cudaSuccess;
  cudaGetLastError();
  mine.ne = h->g.ne; mine.ny = h->g.ny; mine.ok = ok;
  DgPeerRecord* d_all = nullptr;
  std::vector<DgPeerRecord> all(R);
  WB_CUDA(cudaMalloc(&d_all, sizeof(DgPeerRecord) * (R + 1)));
  WB_CUDA(cudaMemcpyAsync(d_all + R, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
  int st = nccl_allgather_bytes(h->comm, d_all + R, d_all, sizeof(DgPeerRecord), h->stream);
  if (st == WB_OK) {
    WB_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(DgPeerRecord) * R, cudaMemcpyDeviceToHost, h->stream));
    WB_CUDA(cudaStreamSynchronize(h->stream));
  }
  cudaFree(d_all);
  WB_CHECK(st);
  for (int k = 0; k < R; ++k) ok = ok && all[k].ok;
  auto open = [&](const cudaIpcMemHandle_t& hd, void** out) {
    return cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
  };
  P.same = (lo_rank >= 0 && lo_rank == hi_rank);
  if (ok && lo_rank >= 0) {
    for (int k = 0; k < 4 && ok; ++k) ok = open(all[lo_rank].buf[k], (void**)&P.lo[k]);
    if (ok) ok = open(all[lo_rank].flags, (void**)&P.lo_flags);
    P.lo_ne = all[lo_rank].ne; P.lo_ny = all[lo_rank].ny;
  }
  if (ok && hi_rank >= 0) {
    if (P.same) {      // a handle is opened once per process
      for (int k = 0; k < 4; ++k) P.hi[k] = P.lo[k];
      P.hi_flags = P.lo_flags;
    } else {
      for (int k = 0; k < 4 && ok; ++k) ok = open(all[hi_rank].buf[k], (void**)&P.hi[k]);
      if (ok) ok = open(all[hi_rank].flags, (void**)&P.hi_flags);
    }
    P.hi_ne = all[hi_rank].ne; P.hi_ny = all[hi_rank].ny;
  }
  cudaGetLastError();
  double bad = ok ? 0.0 : 1.0;      // every rank has its neighbours mapped, or nobody uses the mappings
  WB_CUDA(cudaMemcpyAsync(h->red, &bad, sizeof(bad), cudaMemcpyHostToDevice, h->stream));
  WB_CHECK(nccl_allreduce_max_f64(h->comm, h->red, 1, h->stream));
  WB_CUDA(cudaMemcpyAsync(&bad, h->red, sizeof(bad), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (bad != 0.0) { dg_peer_close(h); return WB_OK; }
  P.active = true;
  P.seq = 0;
  return WB_OK;
}

// one fused launch: out = limiter(c0*A0 + c1*A1 + cd*dt*L(in)) [+ the optional second combination]
int dg_stage_fast(wb_dg2d* h, const double* in, double* out, const double* A0, double c0, const double* A1, double c1, double cd,
                  double* out2 = nullptr, const double* B0 = nullptr, double k0 = 0, const double* B1 = nullptr, double k1 = 0,
                  double k2 = 0, double k3 = 0, double ke = 0) {
  const bool lim_after = h->prm.limiter_id >= 2;
  double* stage_out = out;
  if (lim_after) {                        // un-limited stage result -> scratch D; k3*w4 of the second result needs the LIMITED w4
    WB_CHECK(dg_ensure(h, &h->D));
    stage_out = h->D;
  }
  StageCoef C;
  C.A0 = A0; C.A1 = A1; C.c0 = c0; C.c1 = c1; C.cd = cd; C.na = A1 ? 2 : 1;
  C.out2 = out2; C.B0 = B0; C.B1 = B1; C.k0 = k0; C.k1 = k1; C.k2 = k2; C.k3 = lim_after ? 0.0 : k3; C.ke = ke;
  C.peer_lo = C.peer_hi = C.peer2_lo = C.peer2_hi = nullptr; C.peer_lo_ne = C.peer_hi_ne = 0;
  const int onp = (h->prm.limiter_id == 1 && h->g.m > 1) ? 1 : 0;
  const CUtensorMap* m_in = nullptr;
  if (h->tma_ok)
    for (int k = 0; k < 4; ++k)
      if (h->map_ptr[k] == in) m_in = &h->map[k];
  if (m_in && h->split_ok && !lim_after && h->nranks > 1 && h->overlap && h->comm_stream && h->g.ny >= 6) {
    // slab with neighbours: first and last owned row first, on the comm stream, followed by their exchange; the interior
    // rows meanwhile on the main stream.  The ghost rows themselves are not computed: the exchange overwrites them.
    const unsigned char* fz = h->phys.ninit == 12 ? h->fz : nullptr;
    const int ny = h->g.ny;
    WB_CUDA(cudaEventRecord(h->ev_main, h->stream));
    WB_CUDA(cudaStreamWaitEvent(h->comm_stream, h->ev_main, 0));
    StageCoef Ce = C;                       // the two boundary-row launches: with the peer ghost rows when they are mapped
    const bool p2p = h->peer.active && dg_peer_targets(h, out, out2, Ce);
    WB_CHECK(launch_stage_split(m_in, in, Ce, stage_out, h->gx, h->gy, fz, h->g, h->phys, h->FB, h->ctrl, onp, h->march_rows, 1, 2, h->comm_stream));
    WB_CHECK(launch_stage_split(m_in, in, Ce, stage_out, h->gx, h->gy, fz, h->g, h->phys, h->FB, h->ctrl, onp, h->march_rows, ny - 2, ny - 1,
                                h->comm_stream));
    if (p2p) {
      wb_dg2d::Peer& Pp = h->peer;
      ++Pp.seq;
      // I am the slab ABOVE my lower neighbour (its word [1]) and BELOW my upper one (its word [0])
      WB_CHECK(peer_signal(Pp.lo_flags ? Pp.lo_flags + 1 : nullptr, Pp.hi_flags ? Pp.hi_flags + 0 : nullptr, Pp.seq, h->comm_stream));
      WB_CHECK(peer_wait(Pp.flags, Pp.lo_flags != nullptr, Pp.hi_flags != nullptr, Pp.seq, h->comm_stream));
    } else {
      WB_CHECK(dg_exchange(h, out, h->comm_stream));
      if (out2) WB_CHECK(dg_exchange(h, out2, h->comm_stream));
    }
    WB_CUDA(cudaEventRecord(h->ev_comm, h->comm_stream));
    WB_CHECK(launch_stage_split(m_in, in, C, stage_out, h->gx, h->gy, fz, h->g, h->phys, h->FB, h->ctrl, onp, h->march_rows, 2, ny - 2, h->stream));
    WB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_comm, 0));
    return WB_OK;
  }
  if (m_in && h->split_ok) {
    WB_CHECK(launch_stage_split(m_in, in, C, stage_out, h->gx, h->gy, h->phys.ninit == 12 ? h->fz : nullptr, h->g, h->phys, h->FB, h->ctrl,
                                onp, h->march_rows, 0, h->g.ny, h->stream));
    wb::g_launches.fetch_sub(1);      // counted again by the WB_LAUNCH_CHECK below
  } else {
    dim3 b(64), gr = elem_grid(h, 64);
    if (h->phys.flux_id >= 2) {
      DISPATCH_M(h, (k_dg_stage_fast<MM, true><<<gr, b, 0, h->stream>>>(in, C, stage_out, h->gx, h->gy, h->phys.ninit == 12 ? h->fz : nullptr,
                                                                      h->g, h->phys, h->FB, h->ctrl, onp)));
    } else {
      DISPATCH_M(h, (k_dg_stage_fast<MM, false><<<gr, b, 0, h->stream>>>(in, C, stage_out, h->gx, h->gy, h->phys.ninit == 12 ? h->fz : nullptr,
                                                                       h->g, h->phys, h->FB, h->ctrl, onp)));
    }
  }
  WB_LAUNCH_CHECK();
  if (lim_after) {      // neighbour-reading limiter: ghost rows of the un-limited result, limiter, ghost rows of the limited one
    WB_CHECK(dg_exchange(h, stage_out));
    WB_CHECK(dg_limit_into(h, stage_out, out, out2, k3));
  }
  WB_CHECK(dg_exchange(h, out));
  if (out2) WB_CHECK(dg_exchange(h, out2));
  return WB_OK;
}

// the fused flow: the stage kernel applies an element-local limiter itself ('ONP' or none); for the neighbour-reading ones
// ('HIO', '1OR', 'LOW', 'POS') it writes the un-limited result to a scratch field and a limiter kernel follows
bool dg_use_fused(const wb_dg2d* h) { return h->arith == 0; }

int dg_step_fused(wb_dg2d* h) {
  double *&du = h->du, *&A = h->A, *Bf = h->Bf, *C = h->C;
  WB_CHECK(dg_max_speed(h, du, 1));
  const int solver = h->prm.solver_id;
  if (solver == 3) {                          // 'EQL' :672-681
    WB_CHECK(dg_stage_fast(h, du, A, du, 1.0, nullptr, 0, 1.0));
    WB_CHECK(dg_stage_fast(h, A, Bf, du, 0.5, A, 0.5, 0.5));
    std::swap(h->du, h->Bf);
  } else if (solver == 1 || solver == 2) {    // SSPRK(5,4) :683-710; 5 launches, L(w3) evaluated once
    WB_CHECK(dg_stage_fast(h, du, A, du, 1.0, nullptr, 0, F32(0.391752226571890)));                                   // w1 -> A
    WB_CHECK(dg_stage_fast(h, A, Bf, du, F32(0.444370493651235), A, F32(0.555629506348765), F32(0.368410593050371))); // w2 -> Bf
    WB_CHECK(dg_stage_fast(h, Bf, A, du, F32(0.620101851488403), Bf, F32(0.379898148511597), F32(0.251891774271694))); // w3 -> A
    WB_CHECK(dg_stage_fast(h, A, C, du, F32(0.178079954393132), A, F32(0.821920045606868), F32(0.544974750228521),    // w4 -> C
                           Bf, du, F32(0.00683325884039), Bf, F32(0.51723167208978), F32(0.12759831133288),           // w5 -> Bf
                           F32(0.34833675773694), F32(0.08460416338212)));
    WB_CHECK(dg_stage_fast(h, C, A, Bf, 1.0, nullptr, 0, F32(0.22600748319395)));                                     // new delta_u -> A
    std::swap(h->du, h->A);
  } else {                                    // 'DEB' :737-747
    WB_CHECK(dg_stage_fast(h, du, A, du, 1.0, nullptr, 0, 1.0));
    std::swap(h->du, h->A);
  }
  k_dg_advance<<<1, 1, 0, h->stream>>>(h->ctrl);
  WB_LAUNCH_CHECK();
  return WB_OK;
}

// one time step of evolve (:666-757)
int dg_step(wb_dg2d* h) {
  if (dg_use_fused(h)) return dg_step_fused(h);
  WB_CHECK(dg_ensure(h, &h->D));
  double *du = h->du, *A = h->A, *Bf = h->Bf, *C = h->C, *D = h->D;
  WB_CHECK(dg_max_speed(h, du, 1));
  const int solver = h->prm.solver_id;
  if (solver == 3) {            // 'EQL' :672-681
    WB_CHECK(dg_update(h, du, D, true));
    WB_CHECK(dg_axpy(h, 1, A, du, 1.0, nullptr, 0, nullptr, 0, nullptr, 0, D, 1.0));
    WB_CHECK(dg_limiter(h, A, true));
    WB_CHECK(dg_update(h, A, D, true));
    WB_CHECK(dg_axpy(h, 2, du, du, 0.5, A, 0.5, nullptr, 0, nullptr, 0, D, 0.5));
    WB_CHECK(dg_limiter(h, du, true));
  } else if (solver == 1 || solver == 2) {   // 'RK4' :683-710 / 'SS4' :711-735 (identical real(4) coefficients)
    WB_CHECK(dg_update(h, du, D, true));
    WB_CHECK(dg_axpy(h, 1, A, du, 1.0, nullptr, 0, nullptr, 0, nullptr, 0, D, F32(0.391752226571890)));          // w1
    WB_CHECK(dg_limiter(h, A, true));
    WB_CHECK(dg_update(h, A, D, true));
    WB_CHECK(dg_axpy(h, 2, Bf, du, F32(0.444370493651235), A, F32(0.555629506348765), nullptr, 0, nullptr, 0, D,
                     F32(0.368410593050371)));                                                                    // w2
    WB_CHECK(dg_limiter(h, Bf, true));
    WB_CHECK(dg_update(h, Bf, D, true));
    WB_CHECK(dg_axpy(h, 2, A, du, F32(0.620101851488403), Bf, F32(0.379898148511597), nullptr, 0, nullptr, 0, D,
                     F32(0.251891774271694)));                                                                    // w3 (over w1)
    WB_CHECK(dg_limiter(h, A, true));
    WB_CHECK(dg_update(h, A, D, true));
    WB_CHECK(dg_axpy(h, 2, C, du, F32(0.178079954393132), A, F32(0.821920045606868), nullptr, 0, nullptr, 0, D,
                     F32(0.544974750228521)));                                                                    // w4
    WB_CHECK(dg_limiter(h, C, true));
    // :700 recomputes compute_update(w3): D still holds exactly that result, so the launch is not repeated
    WB_CHECK(dg_axpy(h, 4, Bf, du, F32(0.00683325884039), Bf, F32(0.51723167208978), A, F32(0.12759831133288), C,
                     F32(0.34833675773694), D, F32(0.08460416338212)));                                           // w5 (over w2)
    WB_CHECK(dg_update(h, C, D, true));
    WB_CHECK(dg_axpy(h, 1, du, Bf, 1.0, nullptr, 0, nullptr, 0, nullptr, 0, D, F32(0.22600748319395)));
    WB_CHECK(dg_limiter(h, du, true));
  } else {                      // 'DEB' :737-747
    WB_CHECK(dg_update(h, du, D, true));
    WB_CHECK(dg_axpy(h, 1, du, du, 1.0, nullptr, 0, nullptr, 0, nullptr, 0, D, 1.0));
    WB_CHECK(dg_limiter(h, du, true));
  }
  k_dg_advance<<<1, 1, 0, h->stream>>>(h->ctrl);
  WB_LAUNCH_CHECK();
  return WB_OK;
}

}  // namespace

extern "C" {

int wb_dg2d_create(wb_dg2d** out, const wb_dg2d_params* p) {
  if (!out || !p) { set_error("null argument"); return WB_ERR_ARG; }
  *out = nullptr;
  WB_REQUIRE(p->nvar == 4, "nvar must be 4 (got %d)", p->nvar);
  WB_REQUIRE(p->nx >= 1 && p->ny >= 1, "nx, ny must be >= 1");
  WB_REQUIRE(p->nx == p->ny, "nx == ny required: the reference wraps x-face neighbours with ny (2d/benchmark_2d_dg.f90:1338)");
  WB_REQUIRE(p->mx == p->my, "mx == my required: the reference integrates the x edges with the x rule in y (:1378)");
  WB_REQUIRE(p->mx >= 1 && p->mx <= MAXM, "mx must be 1..%d (got %d)", MAXM, p->mx);
  WB_REQUIRE(p->bc >= 1 && p->bc <= 3, "bc must be 1..3");
  WB_REQUIRE(p->source >= 1 && p->source <= 3, "source must be 1..3");
  WB_REQUIRE(p->grad_phi_case == 1 || p->grad_phi_case == 2, "grad_phi_case must be 1 or 2");
  WB_REQUIRE(p->flux_id >= 0 && p->flux_id <= 3, "flux_id must be 0 (as shipped), 1 (llf1), 2 (hll2) or 3 (hllc)");
  WB_REQUIRE(p->limiter_id >= 0 && p->limiter_id <= 6, "limiter_id must be 0..6 (none, ONP, HIO, 1OR, LOW, POS, PO3)");
  WB_REQUIRE(p->solver_id >= 1 && p->solver_id <= 4, "solver_id must be 1..4 (RK4, SS4, EQL, DEB)");
  WB_REQUIRE(p->gamma > 1.0 && p->boxlen_x > 0 && p->boxlen_y > 0 && p->cfl > 0, "gamma>1, boxlen>0, cfl>0 required");
  WB_REQUIRE(p->arith == 0 || p->arith == 1, "arith must be 0 (fused) or 1 (reference order)");
  const int nranks = p->nranks <= 0 ? 1 : p->nranks;      // 0 (zero-initialised struct) means "no slabs"
  WB_REQUIRE(p->rank >= 0 && p->rank < nranks, "bad rank/nranks %d/%d", p->rank, p->nranks);
  WB_REQUIRE(nranks == 1 || p->arith == 0, "slab mode (nranks > 1) is built for the fused flow: arith 0");
  WB_REQUIRE(nranks == 1 || p->limiter_id != 6 || p->bc == 1, "limiter 'PO3' wraps periodically whatever bc is: on slabs only with bc = 1 (ring of ranks)");
  WB_REQUIRE(nranks == 1 || p->ny / nranks >= 1, "each slab needs at least one row");
  int dev = 0;
  WB_CHECK(select_device(p->device, &dev));
  wb_dg2d* h = new wb_dg2d;
  h->prm = *p;
  h->dev = dev;
  h->rank = p->rank; h->nranks = nranks;
  DgGrid& g = h->g;
  g.nx = p->nx; g.m = p->mx; g.nm = p->mx * p->my; g.nyg = p->ny;
  if (nranks > 1) {        // y slabs: rank r owns global rows [ny*r/R, ny*(r+1)/R) plus one ghost row on each side
    const int ja = (int)((long long)p->ny * p->rank / nranks), jb = (int)((long long)p->ny * (p->rank + 1) / nranks);
    h->nyl = jb - ja;
    g.ny = h->nyl + 2; g.j0 = ja - 1; g.slab = 1; g.e_off = (size_t)p->nx; g.ne_own = (size_t)p->nx * h->nyl;
  } else {
    h->nyl = p->ny;
    g.ny = p->ny; g.j0 = 0; g.slab = 0; g.e_off = 0; g.ne_own = (size_t)p->nx * p->ny;
  }
  g.ne = (size_t)g.nx * g.ny;
  h->nfield = (size_t)4 * g.nm * g.ne;
  h->B = make_basis(p->mx);
  {
    std::memset(&h->FB, 0, sizeof(h->FB));
    const Basis& B0 = h->B;
    for (int q = 0; q < MAXM; ++q)
      for (int n = 0; n < MAXM; ++n) {
        h->FB.P[q][n] = B0.P[q][n]; h->FB.Pw[q][n] = B0.P[q][n] * B0.wq[q]; h->FB.dPw[q][n] = B0.dP[q][n] * B0.wq[q];
        h->FB.Pg[q][n] = B0.Pg[q][n];
        h->FB.EpEp[q][n] = B0.Ep[q] * B0.Ep[n];
        h->FB.Pwh[q][n] = 0.5 * h->FB.Pw[q][n];
      }
    for (int n = 0; n < MAXM; ++n) { h->FB.Em[n] = B0.Em[n]; h->FB.Ep[n] = B0.Ep[n]; }
    // odd number of Gauss points: the middle node is the origin (the reference's Newton iterate is 0 or ~1e-17).  In the
    // fused tables the odd polynomials and the derivatives of the even ones are exactly 0 there (dg2d_fast.cuh, zP / zD)
    if (p->mx & 1) {
      const int qm = p->mx / 2;
      if (std::fabs(B0.xq[qm]) > 1e-14) { set_error("basis tables: middle Gauss node is not the origin"); delete h; return WB_ERR_STATE; }
      for (int n = 0; n < MAXM; ++n) {
        if (n & 1) { h->FB.P[qm][n] = 0.0; h->FB.Pw[qm][n] = 0.0; h->FB.Pwh[qm][n] = 0.0; }
        else h->FB.dPw[qm][n] = 0.0;
      }
    }
    // the fused kernels drop the terms these identities make trivial (dg2d_fast.cuh, trace1)
    for (int q = 0; q < p->mx; ++q)
      if (B0.Em[0] != 1.0 || B0.Ep[0] != 1.0 || B0.P[q][0] != 1.0 || h->FB.dPw[q][0] != 0.0) {
        set_error("basis tables: P_0 is not identically 1"); delete h; return WB_ERR_STATE;
      }
    h->FB.gll = B0.gll;
  }
  h->arith = p->arith;
  DgPhys& P = h->phys;
  P.gamma = p->gamma; P.gm1a = p->gamma - (double)1.0f; P.gm1b = p->gamma - (double)1.f;
  P.dx = p->boxlen_x / (double)p->nx;
  P.oneoverdx = 1. / P.dx;
  P.eps = p->eps; P.M = p->M;
  P.rho_floor = (double)10e-10f; P.p_floor = 1e-10;
  P.gsep = 0; P.grow = 0;
  P.bc = p->bc; P.source = p->source; P.flux_id = p->flux_id; P.ninit = p->ninit;
  {
    const int gll = h->B.gll;
    double gll_w_1 = (p->mx == 1) ? 1. : 1. / (double)((float)(gll * (gll - 1)) + 1e-10f);   // :651-655
    P.dt_num = p->cfl * std::fmin(1. / (double)(2 * 4 + 1), gll_w_1 / 2.);                    // :671
  }
  auto fail = [&](int s) { wb_dg2d_destroy(h); return s; };
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(WB_ERR_CUDA); }
  h->own_stream = true;
  const size_t fb = sizeof(double) * h->nfield, nb = sizeof(double) * g.nm * g.ne;
  h->nparts = (int)std::min<size_t>((g.ne + 255) / 256, 148 * 8);
  cudaError_t e = cudaSuccess;
  // dudt (D), limiter scratch (E) and the host staging buffer are allocated on first use: the fused flow needs none of them
  double** bufs[] = {&h->du, &h->A, &h->Bf, &h->C};
  for (double** b : bufs)
    if (e == cudaSuccess) e = cudaMalloc(b, fb);
  const bool need_xy = (p->source == 2) || (p->ninit == 12);     // grad_phi / freeze mask need the node coordinates
  if (need_xy) {
    if (e == cudaSuccess) e = cudaMalloc(&h->gx, nb);
    if (e == cudaSuccess) e = cudaMalloc(&h->gy, nb);
    if (e == cudaSuccess) e = cudaMalloc(&h->xy, 2 * nb);
    if (e == cudaSuccess) e = cudaMalloc(&h->fz, (size_t)g.nm * g.ne);
  }
  if (e == cudaSuccess) e = cudaMalloc(&h->ctrl, sizeof(DgCtrl));
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_ctrl, sizeof(DgCtrl));
  if (e == cudaSuccess) e = cudaMalloc(&h->part1, sizeof(SpeedKey) * h->nparts);
  if (e == cudaSuccess) e = cudaMalloc(&h->part2, sizeof(double) * h->nparts);
  if (g.slab) {
    const size_t rb = sizeof(double) * 4 * g.nm * g.nx;
    double** rows[] = {&h->sbuf_lo, &h->sbuf_hi, &h->rbuf_lo, &h->rbuf_hi};
    for (double** b : rows)
      if (e == cudaSuccess) e = cudaMalloc(b, rb);
    if (e == cudaSuccess) e = cudaMalloc(&h->red, sizeof(double) * 8);
    if (h->nranks > 1) {
      int prio_lo = 0, prio_hi = 0;
      cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
      if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, prio_hi);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming);
      if (const char* ev = getenv("WB_DG2D_OVERLAP")) h->overlap = atoi(ev) != 0;
    }
  }
  if (e != cudaSuccess) { set_error("device allocation failed: %s", cudaGetErrorString(e)); return fail(WB_ERR_CUDA); }
  cudaMemsetAsync(h->ctrl, 0, sizeof(DgCtrl), h->stream);
  if (need_xy) {
    cudaMemsetAsync(h->gx, 0, nb, h->stream);
    cudaMemsetAsync(h->gy, 0, nb, h->stream);
    cudaMemsetAsync(h->fz, 0, (size_t)g.nm * g.ne, h->stream);
  }
  if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) { set_error("init failed: %s", cudaGetErrorString(e)); return fail(WB_ERR_CUDA); }
  {      // k_dg_stage_split: TMA-staged rows, blocks of 32 elements of a row x 4 variables (WB_DG2D_TMA=0: global-memory kernel)
    const char* env = getenv("WB_DG2D_TMA");
    const char* envr = getenv("WB_DG2D_ROWS");
    if (envr && atoi(envr) > 0) h->march_rows = atoi(envr);
    // a TMA box may be wider than the tensor (nx = 32 < DGT_W = 36: the columns outside are zero-filled like any other
    // out-of-range column; measured bit-identical to the global-memory path on B200), so 32 is the smallest grid
    if (p->arith == 0 && g.nx % 32 == 0 && g.nx >= 32 && !(env && atoi(env) == 0)) {
      const double* bufs4[4] = {h->du, h->A, h->Bf, h->C};
      for (int k = 0; k < 4; ++k) {
        int st = dg_make_map(h, bufs4[k], &h->map[k]);
        if (st != WB_OK) return fail(st);
        h->map_ptr[k] = bufs4[k];
      }
      h->tma_ok = true;
      h->split_ok = true;
    }
  }
  *out = h;
  return WB_OK;
}

int wb_dg2d_destroy(wb_dg2d* h) {
  if (!h) return WB_OK;
  cudaSetDevice(h->dev);
  if (h->stream) cudaStreamSynchronize(h->stream);
  output_wait(&h->out_job);
  cudaFree(h->du); cudaFree(h->A); cudaFree(h->Bf); cudaFree(h->C); cudaFree(h->D); cudaFree(h->E); cudaFree(h->stage);
  cudaFree(h->gx); cudaFree(h->gy); cudaFree(h->xy); cudaFree(h->fz); cudaFree(h->ctrl); cudaFree(h->part1); cudaFree(h->part2);
  cudaFree(h->sbuf_lo); cudaFree(h->sbuf_hi); cudaFree(h->rbuf_lo); cudaFree(h->rbuf_hi); cudaFree(h->red);
  if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
  if (h->peer.active && h->comm) {      // nobody frees a buffer a neighbour still has mapped: close, meet, then free
    dg_peer_close(h);
    nccl_allreduce_max_f64(h->comm, h->red, 1, h->stream);
    cudaStreamSynchronize(h->stream);
  }
  cudaFree(h->peer.flags);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  if (h->ev_main) cudaEventDestroy(h->ev_main);
  if (h->ev_comm) cudaEventDestroy(h->ev_comm);
  nccl_comm_destroy(h->comm);
  if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return WB_OK;
}

int wb_dg2d_set_stream(wb_dg2d* h, void* cuda_stream) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  h->stream = (cudaStream_t)cuda_stream;
  return WB_OK;
}

int wb_dg2d_comm_init(wb_dg2d* h, const void* id128) {
  if (!h || !id128) { set_error("null argument"); return WB_ERR_ARG; }
  WB_REQUIRE(h->nranks > 1, "comm_init needs nranks > 1");
  WB_CUDA(cudaSetDevice(h->dev));
  if (h->comm) { dg_peer_close(h); nccl_comm_destroy(h->comm); h->comm = nullptr; }
  WB_CHECK(nccl_comm_create(&h->comm, id128, h->rank, h->nranks));
  return dg_peer_setup(h);
}

const char* wb_dg2d_exchange_kind(const wb_dg2d* h) {
  if (!h || h->nranks <= 1) return "none";
  return h->peer.active ? "p2p" : "nccl";
}

int wb_dg2d_local_rows(const wb_dg2d* h, int* j0, int* nrows) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  if (j0) *j0 = h->g.slab ? h->g.j0 + 1 : 0;
  if (nrows) *nrows = h->nyl;
  return WB_OK;
}

const char* wb_dg2d_stage_kernel(const wb_dg2d* h) {
  if (!h) return "";
  if (!dg_use_fused(h)) return "reference";
  if (h->tma_ok && h->split_ok) return "split";
  return "fast";
}

int wb_dg2d_quadrature(wb_dg2d* h, double* x_quad, double* w_quad) {
  if (!h || !x_quad || !w_quad) { set_error("null argument"); return WB_ERR_ARG; }
  for (int i = 0; i < h->g.m; ++i) { x_quad[i] = h->B.xq[i]; w_quad[i] = h->B.wq[i]; }
  return WB_OK;
}

int wb_dg2d_get_modes_from_nodes(wb_dg2d* h, const double* nodes, double* modes) {
  if (!h || !nodes || !modes) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  h->resident = false;
  WB_CHECK(dg_h2d_field(h, nodes, h->A));
  dim3 b(128), gr = elem_grid(h, 128);
  DISPATCH_M(h, k_modes_from_nodes<MM><<<gr, b, 0, h->stream>>>(h->A, h->Bf, h->g, h->B));
  WB_LAUNCH_CHECK();
  return dg_d2h_field(h, h->Bf, modes);
}

static int dg_fill_initial_nodes(wb_dg2d* h, int ninit, double eta, bool want_xy, double shift_x, double shift_y);

// compute_error :23-89 of the nodal fields in buffers a and b
static int dg_error_of(wb_dg2d* h, const double* a, const double* b, double* lmax4, double* l1_4, double* l2_4) {
  const int nb = (int)std::min<size_t>((h->g.ne_own + 127) / 128, 148 * 4);
  double* part = nullptr;
  WB_CUDA(cudaMalloc(&part, sizeof(double) * 12 * (nb + 1)));
  const double scale = (1.0 / (double)h->prm.nx) * (1.0 / (double)h->prm.ny);      // dx*dy with dx = 1./dble(nx) (:35-36)
  DISPATCH_M(h, k_dg_error<MM><<<nb, 128, 0, h->stream>>>(a, b, h->g, h->B, scale, part));
  k_dg_error_final<<<1, 32, 0, h->stream>>>(part, nb, part + (size_t)12 * nb);
  wb::g_launches.fetch_add(2);
  double r[12];
  cudaError_t e1 = cudaMemcpyAsync(r, part + (size_t)12 * nb, sizeof(r), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e2 = cudaStreamSynchronize(h->stream);
  cudaFree(part);
  if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("compute_error: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2)); return WB_ERR_CUDA; }
  for (int k = 0; k < 4; ++k) { lmax4[k] = r[k]; l1_4[k] = r[4 + k]; l2_4[k] = r[8 + k]; }
  return WB_OK;
}

int wb_dg2d_compute_error(wb_dg2d* h, const double* u_nodes, const double* u_init_nodes, double* lmax4, double* l1_4, double* l2_4) {
  if (!h || !u_nodes || !u_init_nodes || !lmax4 || !l1_4 || !l2_4) { set_error("null argument"); return WB_ERR_ARG; }
  if (h->nranks > 1) { set_error("wb_dg2d_compute_error: single-GPU handles only"); return WB_ERR_STATE; }
  WB_CUDA(cudaSetDevice(h->dev));
  h->resident = false;
  WB_CHECK(dg_h2d_field(h, u_nodes, h->A));
  WB_CHECK(dg_h2d_field(h, u_init_nodes, h->Bf));
  return dg_error_of(h, h->A, h->Bf, lmax4, l1_4, l2_4);
}

int wb_dg2d_compute_error_resident(wb_dg2d* h, int ninit, double eta, double shift_x, double shift_y, double* lmax4, double* l1_4,
                                   double* l2_4) {
  if (!h || !lmax4 || !l1_4 || !l2_4) { set_error("null argument"); return WB_ERR_ARG; }
  if (!h->resident) { set_error("no resident state"); return WB_ERR_STATE; }
  if (h->nranks > 1) { set_error("wb_dg2d_compute_error_resident: single-GPU handles only"); return WB_ERR_STATE; }
  WB_REQUIRE(ninit >= 1 && ninit <= 12, "ninit must be 1..12 (got %d)", ninit);
  WB_CUDA(cudaSetDevice(h->dev));
  WB_CHECK(dg_fill_initial_nodes(h, ninit, eta, false, shift_x, shift_y));                        // u_anal -> A
  dim3 b(128), gr = elem_grid(h, 128);
  DISPATCH_M(h, k_nodes_from_modes<MM><<<gr, b, 0, h->stream>>>(h->du, h->Bf, h->g, h->B));      // nodes of the state -> Bf
  WB_LAUNCH_CHECK();
  return dg_error_of(h, h->Bf, h->A, lmax4, l1_4, l2_4);
}

int wb_dg2d_get_nodes_from_modes(wb_dg2d* h, const double* modes, double* nodes) {
  if (!h || !nodes || !modes) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  h->resident = false;
  WB_CHECK(dg_h2d_field(h, modes, h->A));
  dim3 b(128), gr = elem_grid(h, 128);
  DISPATCH_M(h, k_nodes_from_modes<MM><<<gr, b, 0, h->stream>>>(h->A, h->Bf, h->g, h->B));
  WB_LAUNCH_CHECK();
  return dg_d2h_field(h, h->Bf, nodes);
}

int wb_dg2d_compute_update(wb_dg2d* h, const double* modes, const double* x, const double* y, double* dudt) {
  if (!h || !modes || !dudt) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  h->resident = false;
  WB_CHECK(dg_set_xy(h, x, y));
  WB_CHECK(dg_ensure(h, &h->D));
  WB_CHECK(dg_h2d_field(h, modes, h->A));
  WB_CHECK(dg_exchange(h, h->A));      // slab handles: the y neighbours of the first / last owned row live on other ranks
  WB_CHECK(dg_update(h, h->A, h->D, false));
  return dg_d2h_field(h, h->D, dudt);
}

int wb_dg2d_apply_limiter(wb_dg2d* h, double* modes_inout) {
  if (!h || !modes_inout) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  h->resident = false;
  k_dg_ctrl_init<<<1, 1, 0, h->stream>>>(h->ctrl, 0.0, -1, 0);
  WB_LAUNCH_CHECK();
  WB_CHECK(dg_h2d_field(h, modes_inout, h->A));
  WB_CHECK(dg_exchange(h, h->A));      // (neighbour-reading limiters on slabs)
  if (h->arith == 0 && h->prm.limiter_id >= 2) {      // the limiter kernels of the fused flow (same operation order, one pass)
    WB_CHECK(dg_limit_into(h, h->A, h->Bf, nullptr, 0.0));
    return dg_d2h_field(h, h->Bf, modes_inout);
  }
  WB_CHECK(dg_limiter(h, h->A, false));
  return dg_d2h_field(h, h->A, modes_inout);
}

int wb_dg2d_compute_max_speed(wb_dg2d* h, const double* mean_mode, double* cs_max, double* v_xmax, double* v_ymax,
                              double* speed_max) {
  if (!h || !mean_mode) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  h->resident = false;
  // mean mode (nvar,nx,ny) -> planes (v, mode 0) of buffer A
  WB_CHECK(dg_ensure(h, &h->stage));
  WB_CUDA(cudaMemcpyAsync(h->stage, mean_mode, sizeof(double) * 4 * h->g.ne_own, cudaMemcpyHostToDevice, h->stream));
  dim3 b(128), gr((unsigned)((h->g.ne_own + 127) / 128), 1);
  k_dg_aos_to_soa<<<gr, b, 0, h->stream>>>(h->stage, h->A, h->g);
  WB_LAUNCH_CHECK();
  WB_CHECK(dg_max_speed(h, h->A, 0));
  WB_CUDA(cudaMemcpyAsync(h->h_ctrl, h->ctrl, sizeof(DgCtrl), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (cs_max) *cs_max = h->h_ctrl->cs_max;
  if (v_xmax) *v_xmax = h->h_ctrl->vx;
  if (v_ymax) *v_ymax = h->h_ctrl->vy;
  if (speed_max) *speed_max = h->h_ctrl->speed_max;
  return WB_OK;
}

int wb_dg2d_upload(wb_dg2d* h, const double* u_nodes, const double* x, const double* y) {
  if (!h || !u_nodes) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  WB_CHECK(dg_set_xy(h, x, y));
  WB_CHECK(dg_h2d_field(h, u_nodes, h->A));
  dim3 b(128), gr = elem_grid(h, 128);
  DISPATCH_M(h, k_modes_from_nodes<MM><<<gr, b, 0, h->stream>>>(h->A, h->du, h->g, h->B));      // :644
  WB_LAUNCH_CHECK();
  k_dg_ctrl_init<<<1, 1, 0, h->stream>>>(h->ctrl, 0.0, -1, 1);
  WB_LAUNCH_CHECK();
  if (h->prm.limiter_id >= 2) WB_CHECK(dg_exchange(h, h->du));      // neighbour-reading limiters need the ghost rows
  WB_CHECK(dg_limiter(h, h->du, false));                                                         // :659
  WB_CHECK(dg_exchange(h, h->du));
  h->resident = true;
  return WB_OK;
}

// nodal initial state of get_initial_conditions (ninit 1..12) into buffer A (and the node coordinates into h->xy)
static int dg_fill_initial_nodes(wb_dg2d* h, int ninit, double eta, bool want_xy, double shift_x, double shift_y) {
  unsigned long long* minbits = reinterpret_cast<unsigned long long*>(h->part2);
  WB_CUDA(cudaMemsetAsync(minbits, 0x7f, sizeof(unsigned long long), h->stream));
  dim3 b(128), gr((unsigned)((h->g.ne + 127) / 128), h->g.nm);
  if (ninit == 1 || ninit == 10 || ninit == 11) {      // w(4) = minval(w(1)) over the whole box
    k_dg_init<<<gr, b, 0, h->stream>>>(h->A, nullptr, h->g, h->phys, h->B, ninit, eta, h->prm.boxlen_x, h->prm.boxlen_y, minbits, 0,
                                       shift_x, shift_y);
    WB_LAUNCH_CHECK();
    if (h->nranks > 1) {       // positive doubles: the bit patterns order like the values
      if (!h->comm) { set_error("nranks > 1 but wb_dg2d_comm_init was not called"); return WB_ERR_STATE; }
      WB_CHECK(nccl_allreduce_min_f64(h->comm, h->part2, 1, h->stream));
    }
  }
  k_dg_init<<<gr, b, 0, h->stream>>>(h->A, want_xy ? h->xy : nullptr, h->g, h->phys, h->B, ninit, eta, h->prm.boxlen_x,
                                     h->prm.boxlen_y, minbits, 1, shift_x, shift_y);
  WB_LAUNCH_CHECK();
  return WB_OK;
}

int wb_dg2d_init_device(wb_dg2d* h, int ninit, double eta) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  WB_REQUIRE(ninit >= 1 && ninit <= 12, "ninit must be 1..12 (got %d)", ninit);
  WB_CUDA(cudaSetDevice(h->dev));
  const size_t n = (size_t)h->g.nm * h->g.ne;
  const bool need_xy = (h->phys.source == 2) || (h->phys.ninit == 12);
  WB_CHECK(dg_fill_initial_nodes(h, ninit, eta, need_xy, 0.0, 0.0));
  if (need_xy) {
    dim3 b2(256), g2((unsigned)((n + 255) / 256));
    if (h->phys.source == 2) {
      k_grad_phi<<<g2, b2, 0, h->stream>>>(h->xy, h->xy + n, h->gx, h->gy, n, h->prm.grad_phi_case);
      WB_LAUNCH_CHECK();
      WB_CHECK(dg_grad_sep(h));
    }
    if (h->phys.ninit == 12) { k_freeze_mask<<<g2, b2, 0, h->stream>>>(h->xy, h->xy + n, h->fz, n, h->prm.boxlen_x / 2., h->prm.boxlen_y / 2.); WB_LAUNCH_CHECK(); }
  }
  h->have_xy = true;
  dim3 be(128), ge = elem_grid(h, 128);
  DISPATCH_M(h, k_modes_from_nodes<MM><<<ge, be, 0, h->stream>>>(h->A, h->du, h->g, h->B));       // :644
  WB_LAUNCH_CHECK();
  k_dg_ctrl_init<<<1, 1, 0, h->stream>>>(h->ctrl, 0.0, -1, 1);
  WB_LAUNCH_CHECK();
  if (h->prm.limiter_id >= 2) WB_CHECK(dg_exchange(h, h->du));      // neighbour-reading limiters need the ghost rows
  WB_CHECK(dg_limiter(h, h->du, false));                                                          // :659
  WB_CHECK(dg_exchange(h, h->du));
  h->resident = true;
  return WB_OK;
}

int wb_dg2d_get_initial_conditions(wb_dg2d* h, int ninit, double eta, double* u_nodes_out) {
  if (!h || !u_nodes_out) { set_error("null argument"); return WB_ERR_ARG; }
  WB_REQUIRE(ninit >= 1 && ninit <= 12, "ninit must be 1..12 (got %d)", ninit);
  WB_CUDA(cudaSetDevice(h->dev));
  h->resident = false;                 // buffer A is scratch of the resident path too
  WB_CHECK(dg_fill_initial_nodes(h, ninit, eta, false, 0.0, 0.0));
  return dg_d2h_field(h, h->A, u_nodes_out);
}

int wb_dg2d_step_async(wb_dg2d* h, int nsteps, double tend) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  if (!h->resident) { set_error("no resident state (call wb_dg2d_upload first)"); return WB_ERR_STATE; }
  WB_CUDA(cudaSetDevice(h->dev));
  k_dg_ctrl_init<<<1, 1, 0, h->stream>>>(h->ctrl, tend, -1, 0);
  WB_LAUNCH_CHECK();
  for (int s = 0; s < nsteps; ++s) WB_CHECK(dg_step(h));
  return WB_OK;
}

int wb_dg2d_sync(wb_dg2d* h, int* iters_out, double* t_out, double* last_dt_out) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  WB_CUDA(cudaMemcpyAsync(h->h_ctrl, h->ctrl, sizeof(DgCtrl), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->peer.active) {
    unsigned long long err = 0;
    WB_CUDA(cudaMemcpy(&err, h->peer.flags + 2, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) { set_error("peer-memory ghost-row exchange %llu timed out (a neighbouring rank stopped)", err); return WB_ERR_NCCL; }
  }
  if (iters_out) *iters_out = h->h_ctrl->iter;
  if (t_out) *t_out = h->h_ctrl->t;
  if (last_dt_out) *last_dt_out = h->h_ctrl->dt;
  return WB_OK;
}

int wb_dg2d_output_file(wb_dg2d* h, int var, int nequilibrium, const char* path) {
  if (!h || !path) { set_error("null argument"); return WB_ERR_ARG; }
  if (!h->resident) { set_error("no resident state"); return WB_ERR_STATE; }
  WB_REQUIRE(var >= 1 && var <= 4, "var must be 1..4 (got %d)", var);
  WB_REQUIRE(nequilibrium >= 1 && nequilibrium <= 3, "nequilibrium must be 1..3 (got %d)", nequilibrium);
  WB_REQUIRE(h->nranks == 1, "output_file: single-GPU handles only (a slab holds part of the table)");
  WB_CUDA(cudaSetDevice(h->dev));
  const int ncol = 2 + (4 - var + 1);
  double* tab = nullptr;
  WB_CUDA(cudaMalloc(&tab, sizeof(double) * ncol * h->g.ne_own));
  dim3 b(128), gr((unsigned)((h->g.ne_own + 127) / 128));
  DISPATCH_M(h, k_dg_pack_output<MM><<<gr, b, 0, h->stream>>>(h->du, h->g, h->phys, h->B, h->prm.boxlen_y, var, nequilibrium, tab));
  WB_LAUNCH_CHECK();
  return output_start(&h->out_job, h->dev, h->stream, tab, h->g.ne_own, ncol, path);
}

int wb_dg2d_output_wait(wb_dg2d* h) {
  if (!h) { set_error("null handle"); return WB_ERR_ARG; }
  return output_wait(&h->out_job);
}

int wb_dg2d_download_modes(wb_dg2d* h, double* modes_out) {
  if (!h || !modes_out) { set_error("null argument"); return WB_ERR_ARG; }
  if (!h->resident) { set_error("no resident state"); return WB_ERR_STATE; }
  WB_CUDA(cudaSetDevice(h->dev));
  return dg_d2h_field(h, h->du, modes_out);
}

int wb_dg2d_download(wb_dg2d* h, double* u_nodes_out) {
  if (!h || !u_nodes_out) { set_error("null argument"); return WB_ERR_ARG; }
  if (!h->resident) { set_error("no resident state"); return WB_ERR_STATE; }
  WB_CUDA(cudaSetDevice(h->dev));
  dim3 b(128), gr = elem_grid(h, 128);
  DISPATCH_M(h, k_nodes_from_modes<MM><<<gr, b, 0, h->stream>>>(h->du, h->A, h->g, h->B));      // :771
  WB_LAUNCH_CHECK();
  return dg_d2h_field(h, h->A, u_nodes_out);
}

int wb_dg2d_evolve(wb_dg2d* h, double* u_nodes_inout, const double* x, const double* y, double tend, int max_iter,
                   int* iters_out, double* t_out, double* last_dt_out) {
  if (!h || !u_nodes_inout) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CHECK(wb_dg2d_upload(h, u_nodes_inout, x, y));
  int iters = 0;
  double t = 0.0, dt = 0.0;
  const int batch = 8;
  for (;;) {
    if (!(t < tend) || (max_iter >= 0 && iters >= max_iter)) break;
    int n = batch;
    if (max_iter >= 0 && max_iter - iters < n) n = max_iter - iters;
    WB_CHECK(wb_dg2d_step_async(h, n, tend));
    WB_CHECK(wb_dg2d_sync(h, &iters, &t, &dt));
  }
  WB_CHECK(wb_dg2d_download(h, u_nodes_inout));
  if (iters_out) *iters_out = iters;
  if (t_out) *t_out = t;
  if (last_dt_out) *last_dt_out = dt;
  return WB_OK;
}

}  // extern "C"
