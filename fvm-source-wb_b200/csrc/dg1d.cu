// 1D modal DG on the perturbation (dg_with_source.f90, default integrator 'RKi'): compute_update_exact_delta
// (:1749-2031), riemann_hllc / riemann_llf (:1318-1374 / :1299-1316), the nodal reconstruction and time-step control
// of the main loop (:173-336).  Basis and quadrature follow the ROOT legendre.f90 (single-precision normalisation
// constants, hard-coded single-precision GL rules for n <= 3).
//
// A correctness configuration of the reference (nx ~ 128): one thread per cell, reference operation order (file
// compiled with -fmad=false), data kept in the reference's u(nvar,n,nx) layout.  Differences to the CPU restatement
// come only from exp() in the face equilibria.
#include "common.cuh"
#include <cmath>
#include <cstring>

namespace wb { namespace dg1d {

constexpr int NV = 3;
constexpr int MAXN = 4;

struct Basis1 { double xq[MAXN], wq[MAXN], P[MAXN][MAXN], dP[MAXN][MAXN], Em[MAXN], Ep[MAXN]; double s05, s6, s10; };
struct P1d { int n, nx, riemann, source, bc, use_limiter; double gamma, boxlen; };
struct CtrlD { double t, dt, tend, cmax; int iter, max_iter, skip; };

// ---- host: root legendre.f90 ------------------------------------------------------------------------------
static double h_legendre(double& x, int n) {                         // :1-25
  x = std::fmin(std::fmax(x, (double)-1.0f), (double)1.0f);
  switch (n) {
    case 0: return (double)(1.0f * sqrtf(0.5f));
    case 1: return x * 0.5 * (double)sqrtf(6.f);
    case 2: return 0.25 * (3.0 * (x * x) - 1.0) * (double)sqrtf(10.f);
    case 3: return 0.5 * (5.0 * ((x * x) * x) - 3.0 * x);
    default: return 0.0;
  }
}
static double h_legendre_prime(double& x, int n) {                   // :27-50
  x = std::fmin(std::fmax(x, (double)-1.0f), (double)1.0f);
  switch (n) {
    case 0: return 0.0;
    case 1: return (double)(1.0f * 0.5f * sqrtf(6.f));
    case 2: return 6.0 * x * 0.25 * (double)sqrtf(10.f);
    case 3: return 0.5 * (15.0 * (x * x) - 3.0);
    default: return 0.0;
  }
}
static void h_gl_quadrature(double* x, double* w, int n) {           // :77-128 (n <= 3: the hard-coded real(4) rules)
  if (n == 1) { x[0] = 0.0; w[0] = 2.0; return; }
  if (n == 2) {
    x[0] = -1.f / (double)3 * (double)sqrtf(3.f); x[1] = 1.f / (double)3 * (double)sqrtf(3.f);
    w[0] = 1.; w[1] = 1.;
    return;
  }
  x[0] = (double)(-sqrtf(3.f) / sqrtf(5.f)); x[1] = 0.0; x[2] = (double)(sqrtf(3.f) / sqrtf(5.f));
  w[0] = (double)(5.f / 9.f); w[1] = (double)(8.f / 9.f); w[2] = (double)(5.f / 9.f);
}
static Basis1 make_basis(int n) {
  Basis1 B;
  std::memset(&B, 0, sizeof(B));
  h_gl_quadrature(B.xq, B.wq, n);
  for (int q = 0; q < n; ++q)
    for (int m = 0; m < n; ++m) {
      double x = B.xq[q]; B.P[q][m] = h_legendre(x, m);
      x = B.xq[q]; B.dP[q][m] = h_legendre_prime(x, m);
    }
  for (int m = 0; m < n; ++m) { double a = -1.0, b = 1.0; B.Em[m] = h_legendre(a, m); B.Ep[m] = h_legendre(b, m); }
  B.s05 = (double)(1.0f * sqrtf(0.5f)); B.s6 = (double)sqrtf(6.f); B.s10 = (double)sqrtf(10.f);   // real(4) constants of legendre.f90
  return B;
}

// ---- device physics (dg_with_source.f90:1167-1246) -----------------------------------------------------------
__device__ __forceinline__ void prim(const double* u, double* w, double gamma) {
  w[0] = u[0];
  w[1] = u[1] / w[0];
  w[2] = (gamma - (double)1.0f) * (u[2] - 0.5 * w[0] * (w[1] * w[1]));
}
__device__ __forceinline__ void cons(const double* w, double* u, double gamma) {
  u[0] = w[0];
  u[1] = w[0] * w[1];
  u[2] = w[2] / (gamma - (double)1.0f) + 0.5 * w[0] * (w[1] * w[1]);
}
__device__ __forceinline__ void flux(const double* u, double* f, double gamma) {
  double w[NV];
  prim(u, w, gamma);
  f[0] = w[1] * u[0];
  f[1] = w[1] * u[1] + w[2];
  f[2] = w[1] * u[2] + w[2] * w[1];
}
__device__ __forceinline__ void source_term(const double* u, double* s, double gamma) {
  double w[NV];
  prim(u, w, gamma);
  s[0] = 0;
  s[1] = -w[0];
  s[2] = -w[0] * w[1];
}
__device__ __forceinline__ double speed(const double* u, double gamma) {
  double w[NV];
  prim(u, w, gamma);
  double cs = sqrt(gamma * fmax(w[2], 1e-10) / fmax(w[0], 1e-10));
  return fabs(w[1]) + cs;
}
__device__ __forceinline__ void riemann_llf(const double* ul, const double* ur, double* fg, double gamma) {   // :1299-1316
  double cl = speed(ul, gamma), cr = speed(ur, gamma), cmax = fmax(cl, cr), fl[NV], fr[NV];
  flux(ul, fl, gamma);
  flux(ur, fr, gamma);
#pragma unroll
  for (int v = 0; v < NV; ++v) fg[v] = 0.5 * (fr[v] + fl[v]) - 0.5 * cmax * (ur[v] - ul[v]);
}
__device__ __forceinline__ void riemann_hllc(const double* ul, const double* ur, double* fg, double gamma) {  // :1318-1374
  double wl[NV], wr[NV];
  prim(ul, wl, gamma);
  prim(ur, wr, gamma);
  double cl = sqrt(gamma * fmax(wl[2], 1e-10) / fmax(wl[0], 1e-10));
  double cr = sqrt(gamma * fmax(wr[2], 1e-10) / fmax(wr[0], 1e-10));
  double SL = fmin(wl[1], wr[1]) - fmax(cl, cr);
  double SR = fmax(wl[1], wr[1]) + fmax(cl, cr);
  double DL = wl[0] * (wl[1] - SL);
  double DR = wr[0] * (SR - wr[1]);
  double ws2 = (DR * wr[1] + DL * wl[1] + (wl[2] - wr[2])) / (DL + DR);
  double ws3 = (DR * wl[2] + DL * wr[2] + DL * DR * (wl[1] - wr[1])) / (DL + DR);
  double wsl1 = wl[0] * (SL - wl[1]) / (SL - ws2);
  double usl3 = ((SL - wl[1]) * ul[2] - wl[2] * wl[1] + ws3 * ws2) / (SL - ws2);
  double wsr1 = wr[0] * (SR - wr[1]) / (SR - ws2);
  double usr3 = ((SR - wr[1]) * ur[2] - wr[2] * wr[1] + ws3 * ws2) / (SR - ws2);
  double g1, g2, g3, e3;
  if (SL > 0.0) { g1 = wl[0]; g2 = wl[1]; g3 = wl[2]; e3 = ul[2]; }
  else if (ws2 > 0.0) { g1 = wsl1; g2 = ws2; g3 = ws3; e3 = usl3; }
  else if (SR > 0.0) { g1 = wsr1; g2 = ws2; g3 = ws3; e3 = usr3; }
  else { g1 = wr[0]; g2 = wr[1]; g3 = wr[2]; e3 = ur[2]; }
  fg[0] = g1 * g2;
  fg[1] = g1 * g2 * g2 + g3;
  fg[2] = g2 * (e3 + g3);
}

#define M3(a, v, i, c) ((a)[((size_t)(c) * P.n + (i)) * NV + (v)])

// compute_update_exact_delta :1749-2031, one thread per cell
__global__ void k_dg1_update(const double* __restrict__ du, const double* __restrict__ u_eq, double* __restrict__ dudt,
                             P1d P, Basis1 B, const CtrlD* ctrl) {
  if (ctrl && ctrl->skip) return;
  int ic = blockIdx.x * blockDim.x + threadIdx.x;     // 0-based cell
  if (ic >= P.nx) return;
  const int n = P.n, nx = P.nx;
  const double gamma = P.gamma;
  if (ic == 0 || ic == nx - 1) {                      // dudt(:,:,1) = dudt(:,:,nx) = 0   :2028-2029
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v) M3(dudt, v, i, ic) = 0;
    return;
  }
  const double dx = P.boxlen / (double)nx, oneoverdx = 1. / dx;
  // face equilibria at x = ic*dx and (ic+1)*dx (0-based cell ic has faces ic+1 and ic+2 in the 1-based reference)
  double ufe[2][NV], ffe[2][NV];
  for (int k = 0; k < 2; ++k) {
    double xf = (double)(ic + k) * dx;
    double w[NV] = {exp(-xf), 0, exp(-xf)};
    cons(w, ufe[k], gamma);
    flux(ufe[k], ffe[k], gamma);
  }
  // volume terms of the cell
  double fq[MAXN][NV], fqe[MAXN][NV], sq[MAXN][NV], sqe[MAXN][NV];
  for (int j = 0; j < n; ++j) {
    double uq[NV] = {0, 0, 0}, us[NV], ue[NV];
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v) uq[v] = uq[v] + M3(du, v, i, ic) * B.P[j][i];
    for (int v = 0; v < NV; ++v) { ue[v] = M3(u_eq, v, j, ic); us[v] = ue[v] + uq[v]; }
    flux(us, fq[j], gamma);
    flux(ue, fqe[j], gamma);
    source_term(us, sq[j], gamma);
    source_term(ue, sqe[j], gamma);
  }
  // traces: own left/right, left neighbour's right, right neighbour's left
  auto trace = [&](int c, const double* E, double* out) {
    for (int v = 0; v < NV; ++v) out[v] = 0.;
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v) out[v] = out[v] + M3(du, v, i, c) * E[i];
  };
  double dl[NV], dr[NV], dnl[NV], dnr[NV];
  trace(ic, B.Em, dl); trace(ic, B.Ep, dr);
  trace(ic - 1, B.Ep, dnl); trace(ic + 1, B.Em, dnr);
  double u_left[NV], u_right[NV], ul_n[NV], ur_n[NV], F0[NV], F1[NV];
  for (int v = 0; v < NV; ++v) {
    u_left[v] = ufe[0][v] + dl[v];        // u_left(icell)  = u_face_eq(icell)   + trace(-1)
    u_right[v] = ufe[1][v] + dr[v];       // u_right(icell) = u_face_eq(icell+1) + trace(+1)
    ul_n[v] = ufe[0][v] + dnl[v];         // u_right(icell-1)
    ur_n[v] = ufe[1][v] + dnr[v];         // u_left(icell+1)
  }
  if (P.riemann == 1) { riemann_llf(ul_n, u_left, F0, gamma); riemann_llf(u_right, ur_n, F1, gamma); }
  else { riemann_hllc(ul_n, u_left, F0, gamma); riemann_hllc(u_right, ur_n, F1, gamma); }
  for (int i = 0; i < n; ++i)
    for (int v = 0; v < NV; ++v) {
      double fv = 0.0, fve = 0.0, sv = 0.0, sve = 0.0;
      for (int j = 0; j < n; ++j) {
        fv = fv + fq[j][v] * B.dP[j][i] * B.wq[j];
        fve = fve + fqe[j][v] * B.dP[j][i] * B.wq[j];
        if (P.source == 2) {
          sv = sv + sq[j][v] * B.P[j][i] * B.wq[j] * 0.5;
          sve = sve + sqe[j][v] * B.P[j][i] * B.wq[j] * 0.5;
        }
      }
      M3(dudt, v, i, ic) = oneoverdx * fv - oneoverdx * fve - oneoverdx * (F1[v] * B.Ep[i] - F0[v] * B.Em[i])
                           + oneoverdx * (ffe[1][v] * B.Ep[i] - ffe[0][v] * B.Em[i]) + sv - sve;
    }
}

// compute_update :807-1028 on the full state (bc 1..4), one thread per cell; thread i evaluates the cell
// ie = (i == 1 ? 2 : i == nx-1 ? nx : i) because dudt(:,:,1) = dudt(:,:,2) and dudt(:,:,nx-1) = dudt(:,:,nx)
__global__ void k_dg1_update_plain(const double* __restrict__ u, double* __restrict__ dudt, P1d P, Basis1 B, const CtrlD* ctrl) {
  if (ctrl && ctrl->skip) return;
  int it = blockIdx.x * blockDim.x + threadIdx.x + 1;     // 1-based
  if (it > P.nx) return;
  const int n = P.n, nx = P.nx;
  const double gamma = P.gamma;
  const int ic = (it == 1) ? 2 : (it == nx - 1 ? nx : it);
  const double dx = P.boxlen / (double)nx, oneoverdx = 1.0 / dx;
  auto trace = [&](int c, const double* E, double* out) {      // c 1-based
    for (int v = 0; v < NV; ++v) out[v] = 0.0;
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v) out[v] = out[v] + M3(u, v, i, c - 1) * E[i];
  };
  auto face = [&](int iface, double* ff) {
    int ileft = iface - 1, iright = iface;
    if (P.bc == 1) { if (iface == 1) ileft = nx; if (iface == nx + 1) iright = 1; }
    if (P.bc == 2 || P.bc == 3) { if (iface == 1) ileft = 1; if (iface == nx + 1) iright = nx; }
    double a[NV] = {0, 0, 0}, b[NV] = {0, 0, 0};
    const bool inb = (ileft >= 1 && iright <= nx);
    if (ileft >= 1) trace(ileft, B.Ep, a);        // u_right(ileft)
    if (iright <= nx) trace(iright, B.Em, b);     // u_left(iright)
    for (int v = 0; v < NV; ++v) ff[v] = 0.0;
    if (inb) { if (P.riemann == 1) riemann_llf(a, b, ff, gamma); else riemann_hllc(a, b, ff, gamma); }
    if ((P.bc == 3 || P.bc == 4) && iface == 1) {
      double src[NV], t[NV];
      if (P.bc == 3) { for (int v = 0; v < NV; ++v) src[v] = b[v]; } else trace(1, B.Ep, src);       // bc 4: u_right(:,1)
      t[0] = src[0]; t[1] = -src[1]; t[2] = src[2];
      if (P.riemann == 1) riemann_llf(t, b, ff, gamma); else riemann_hllc(t, b, ff, gamma);
    }
    if ((P.bc == 3 || P.bc == 4) && iface == nx + 1) {
      double src[NV], t[NV];
      if (P.bc == 3) { for (int v = 0; v < NV; ++v) src[v] = a[v]; } else trace(nx, B.Em, src);       // bc 4: u_left(:,nx)
      t[0] = src[0]; t[1] = -src[1]; t[2] = src[2];
      if (P.riemann == 1) riemann_llf(a, t, ff, gamma); else riemann_hllc(a, t, ff, gamma);
    }
  };
  double fq[MAXN][NV], sq[MAXN][NV], F0[NV], F1[NV];
  for (int j = 0; j < n; ++j) {
    double uq[NV] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v) uq[v] = uq[v] + M3(u, v, i, ic - 1) * B.P[j][i];
    flux(uq, fq[j], gamma);
    source_term(uq, sq[j], gamma);
  }
  face(ic, F0);
  face(ic + 1, F1);
  for (int i = 0; i < n; ++i)
    for (int v = 0; v < NV; ++v) {
      double fv = 0.0, sv = 0.0;
      for (int j = 0; j < n; ++j) {
        fv = fv + fq[j][v] * B.dP[j][i] * B.wq[j];
        if (P.source == 2) sv = sv + sq[j][v] * B.P[j][i] * B.wq[j];
      }
      M3(dudt, v, i, it - 1) = oneoverdx * (fv - (F1[v] * B.Ep[i] - F0[v] * B.Em[i])) + sv;
    }
}

// cons_to_prim :1225-1233, prim_to_cons :1250-1258, cons_to_char :1260-1272, char_to_cons :1274-1286
__device__ __forceinline__ void cons_to_prim(const double* du, double* dw, const double* w, double gamma) {
  dw[0] = du[0];
  dw[1] = (du[1] - w[1] * du[0]) / w[0];
  dw[2] = (gamma - (double)1.0f) * (0.5 * (w[1] * w[1]) * du[0] - w[1] * du[1] + du[2]);
}
__device__ __forceinline__ void prim_to_cons(const double* dw, double* du, const double* w, double gamma) {
  du[0] = dw[0];
  du[1] = w[1] * dw[0] + w[0] * dw[1];
  du[2] = 0.5 * (w[1] * w[1]) * dw[0] + w[0] * w[1] * dw[1] + dw[2] / (gamma - (double)1.0f);
}
__device__ __forceinline__ void cons_to_char(const double* du, double* dw, const double* w, double gamma) {
  double csq = gamma * fmax(w[2], 1e-10) / fmax(w[0], 1e-10), cs = sqrt(csq), dp[NV];
  cons_to_prim(du, dp, w, gamma);
  dw[0] = dp[0] - dp[2] / csq;
  dw[1] = 0.5 * (dp[2] / csq + dp[1] * w[0] / cs);
  dw[2] = 0.5 * (dp[2] / csq - dp[1] * w[0] / cs);
}
__device__ __forceinline__ void char_to_cons(const double* dw, double* du, const double* w, double gamma) {
  double csq = gamma * fmax(w[2], 1e-10) / fmax(w[0], 1e-10), cs = sqrt(csq), dp[NV];
  dp[0] = dw[0] + dw[1] + dw[2];
  dp[1] = (dw[1] - dw[2]) * cs / w[0];
  dp[2] = (dw[1] + dw[2]) * csq;
  prim_to_cons(dp, du, w, gamma);
}
__device__ __forceinline__ double minmod3(double x, double y, double z) {
  double s = copysign(1.0, x);
  if (copysign(1.0, y) == s && copysign(1.0, z) == s) return s * fmin(fmin(fabs(x), fabs(y)), fabs(z));
  return 0.0;
}
// limiter(u) :414-519: reads `u` (un-limited, with neighbours), writes `ul`
__global__ void k_dg1_limiter(const double* __restrict__ u, double* __restrict__ ul, P1d P, const CtrlD* ctrl) {
  if (ctrl && ctrl->skip) return;
  int ic = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (ic > P.nx) return;
  const int n = P.n, nx = P.nx;
  const double gamma = P.gamma;
  double el[MAXN][NV];
  for (int i = 0; i < n; ++i)
    for (int v = 0; v < NV; ++v) el[i][v] = M3(u, v, i, ic - 1);
  if (P.use_limiter) {
    int ileft = ic - 1, iright = ic + 1;
    double switch_left = 1.0, switch_right = 1.0;
    if (P.bc == 1) { if (ic == 1) ileft = nx; if (ic == nx) iright = 1; }
    if (P.bc == 2) { if (ic == 1) ileft = 1; if (ic == nx) iright = nx; }
    if (P.bc == 3) { if (ic == 1) { ileft = 1; switch_left = -1.0; } if (ic == nx) { iright = nx; switch_right = -1.0; } }
    if (ileft >= 1 && iright <= nx) {
      double w[NV], wL[MAXN][NV], wM[MAXN][NV], wR[MAXN][NV], w_lim[MAXN][NV];
      prim(&M3(u, 0, 0, ic - 1), w, gamma);
      for (int i = n - 1; i >= 1; --i) {
        double coeff_i = sqrt(2.0 * (double)(i - 1) + 1.0) * (2.0 * (double)i - 1);
        double coeff_ip1 = sqrt(2.0 * (double)i + 1.0) * (2.0 * (double)i - 1);
        double uL[NV], uR[NV], uM[NV];
        for (int v = 0; v < NV; ++v) {
          uL[v] = (M3(u, v, i - 1, ic - 1) - M3(u, v, i - 1, ileft - 1)) * coeff_i / coeff_ip1;
          uR[v] = (M3(u, v, i - 1, iright - 1) - M3(u, v, i - 1, ic - 1)) * coeff_i / coeff_ip1;
          uM[v] = M3(u, v, i, ic - 1);
        }
        uL[1] = switch_left * uL[1];
        uR[1] = switch_right * uR[1];
        cons_to_char(uL, wL[i], w, gamma);
        cons_to_char(uR, wR[i], w, gamma);
        cons_to_char(uM, wM[i], w, gamma);
      }
      for (int i = 1; i < n; ++i)
        for (int v = 0; v < NV; ++v) w_lim[i][v] = wM[i][v];
      for (int v = 0; v < NV; ++v)
        for (int i = n - 1; i >= 1; --i) {
          double w_min = minmod3(wL[i][v], wM[i][v], wR[i][v]);
          w_lim[i][v] = w_min;
          if (fabs(w_min - wM[i][v]) < (double)0.01f * fabs(wM[i][v])) break;
        }
      for (int i = n - 1; i >= 1; --i) char_to_cons(w_lim[i], el[i], w, gamma);
    }
  }
  {
    double w[NV], u_left[NV] = {0, 0, 0}, u_right[NV] = {0, 0, 0}, w_left[NV], w_right[NV];
    prim(el[0], w, gamma);
    for (int i = 1; i <= n; ++i)
      for (int v = 0; v < NV; ++v) {
        u_left[v] = u_left[v] + el[i - 1][v] * ((i - 1) % 2 == 0 ? 1.0 : -1.0) * sqrt(2.0 * (double)i - 1.0);
        u_right[v] = u_right[v] + el[i - 1][v] * sqrt(2.0 * (double)i - 1.0);
      }
    cons_to_prim(u_left, w_left, w, gamma);
    cons_to_prim(u_right, w_right, w, gamma);
    if (w_left[0] < 1e-10 || w_right[0] < 1e-10 || w_left[2] < 1e-10 || w_left[2] < 1e-10)
      for (int i = 1; i < n; ++i)
        for (int v = 0; v < NV; ++v) el[i][v] = 0.0;
  }
  for (int i = 0; i < n; ++i)
    for (int v = 0; v < NV; ++v) M3(ul, v, i, ic - 1) = el[i][v];
}

// nodal reconstruction :313-330: legendre at the PHYSICAL coordinate (clamped to [-1,1]), as shipped
__device__ __forceinline__ double d_legendre(const Basis1& B, double& x, int n) {
  x = fmin(fmax(x, (double)-1.0f), (double)1.0f);
  switch (n) {
    case 0: return B.s05;                                     // 1.0*sqrt(0.5) in real(4)
    case 1: return x * 0.5 * B.s6;                            // sqrt(6.) in real(4)
    case 2: return 0.25 * (3.0 * (x * x) - 1.0) * B.s10;      // sqrt(10.) in real(4)
    case 3: return 0.5 * (5.0 * ((x * x) * x) - 3.0 * x);
    default: return 0.0;
  }
}
__global__ void k_dg1_reconstruct(const double* __restrict__ du, const double* __restrict__ u_eq, double* __restrict__ uinit,
                                  P1d P, Basis1 B, const CtrlD* ctrl) {
  if (ctrl && ctrl->skip) return;
  int ic = blockIdx.x * blockDim.x + threadIdx.x;
  if (ic >= P.nx) return;
  const double dx = P.boxlen / (double)P.nx;
  const double xcell = ((double)(ic + 1) - 0.5) * dx;
  for (int i = 0; i < P.n; ++i) {
    double xq = xcell + dx / 2.0 * B.xq[i];
    for (int v = 0; v < NV; ++v) {
      double a = 0.0;
      for (int m = 0; m < P.n; ++m) a = a + M3(du, v, m, ic) * d_legendre(B, xq, m);
      M3(uinit, v, i, ic) = M3(u_eq, v, i, ic) + a;
    }
  }
}
// compute_max_speed :1136-1152 (first node of each cell) and dt = 0.9*dx/cmax/(2n+1) (:178)
__global__ void k_dg1_max_speed(const double* __restrict__ u_nodes, P1d P, CtrlD* ctrl, int set_dt) {
  __shared__ double sh[256];
  double m = 0.0;
  for (int c = threadIdx.x; c < P.nx; c += blockDim.x) m = fmax(m, speed(u_nodes + (size_t)c * P.n * NV, P.gamma));
  sh[threadIdx.x] = m;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ctrl->cmax = sh[0];
    if (set_dt) {
      const bool done = !(ctrl->t < ctrl->tend) || (ctrl->max_iter >= 0 && ctrl->iter >= ctrl->max_iter);
      ctrl->skip = done ? 1 : 0;
      if (!done) ctrl->dt = (double)0.9f * (P.boxlen / (double)P.nx) / sh[0] / (2.0 * (double)P.n + 1.0);
    }
  }
}

// compute_update_exact :1380-1744 (integrator 'RKw'): full-state modes u, equilibrium MODES u_eq; bc 4 | 5 only (any
// other bc uses the out-of-bounds reads u_right(:,0) / u_left(:,nx+1) of the reference).  One thread per cell.
__device__ __forceinline__ void eq_cons(double x, double* u, double gamma) {
  double w[NV] = {exp(-x), 0, exp(-x)};
  cons(w, u, gamma);
}
__global__ void k_dg1_update_exact(const double* __restrict__ u, const double* __restrict__ u_eq, double* __restrict__ dudt, P1d P,
                                   Basis1 B, const CtrlD* ctrl) {
  if (ctrl && ctrl->skip) return;
  int ic = blockIdx.x * blockDim.x + threadIdx.x;     // 0-based cell
  if (ic >= P.nx) return;
  const int n = P.n, nx = P.nx;
  const double gamma = P.gamma;
  const double dx = P.boxlen / (double)nx, oneoverdx = 1. / dx;
  double ufe[2][NV], ffe[2][NV];                       // face equilibria at x = ic*dx and (ic+1)*dx
  for (int k = 0; k < 2; ++k) { eq_cons((double)(ic + k) * dx, ufe[k], gamma); flux(ufe[k], ffe[k], gamma); }
  double fq[MAXN][NV], fqe[MAXN][NV], sq[MAXN][NV], sqe[MAXN][NV];
  for (int j = 0; j < n; ++j) {
    double uq[NV] = {0, 0, 0}, uqe[NV] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v) {
        uq[v] = uq[v] + M3(u, v, i, ic) * B.P[j][i];
        uqe[v] = uqe[v] + M3(u_eq, v, i, ic) * B.P[j][i];
      }
    flux(uq, fq[j], gamma);
    flux(uqe, fqe[j], gamma);
    source_term(uq, sq[j], gamma);
    source_term(uqe, sqe[j], gamma);
  }
  // u_left(c) = u_face_eq(c) + (trace_-(c) - U_eq(x_left(c))),  u_right(c) = u_face_eq(c+1) + (trace_+(c) - U_eq(x_right(c)))
  auto side = [&](int c, int right, double* out) {       // c 0-based
    double t[NV] = {0, 0, 0}, ue[NV], uf[NV];
    const double* E = right ? B.Ep : B.Em;
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v) t[v] = t[v] + M3(u, v, i, c) * E[i];
    eq_cons((double)(c + right) * dx, ue, gamma);
    eq_cons((double)(c + right) * dx, uf, gamma);        // u_face_eq(c + right + 1): the same point, the same arithmetic
    for (int v = 0; v < NV; ++v) out[v] = uf[v] + (t[v] - ue[v]);
  };
  double F0[NV], F1[NV], a[NV], b[NV], t[NV], s0[NV], s1[NV];
  // left face of the cell (1-based face ic+1)
  if (ic == 0) {
    if (P.bc == 4) {
      eq_cons((double)-0.5f * dx + dx / 2.0 * (double)(1), a, gamma);
      eq_cons((double)0.5f * dx + dx / 2.0 * (double)(-1), b, gamma);
      side(0, 0, s0);
      for (int v = 0; v < NV; ++v) t[v] = a[v] + s0[v] - b[v];
    } else {
      eq_cons((double)-0.5f * dx, a, gamma);
      eq_cons((double)0.5f * dx, b, gamma);
      for (int v = 0; v < NV; ++v) t[v] = a[v] + M3(u, v, 0, 0) - b[v];
    }
    flux(t, F0, gamma);
  } else {
    side(ic - 1, 1, s0); side(ic, 0, s1);
    if (P.riemann == 1) riemann_llf(s0, s1, F0, gamma); else riemann_hllc(s0, s1, F0, gamma);
  }
  // right face (1-based face ic+2)
  if (ic == nx - 1) {
    if (P.bc == 4) {
      eq_cons((double)((float)nx + 0.5f) * dx + dx / 2.0 * (double)(-1), a, gamma);
      eq_cons((double)nx * dx, b, gamma);
      side(nx - 1, 1, s0);
      for (int v = 0; v < NV; ++v) t[v] = a[v] + s0[v] - b[v];
    } else {
      eq_cons((double)((float)nx + 0.5f) * dx, a, gamma);
      eq_cons((double)((float)nx - 0.5f) * dx, b, gamma);
      for (int v = 0; v < NV; ++v) t[v] = a[v] + M3(u, v, 0, nx - 1) - b[v];
    }
    flux(t, F1, gamma);
  } else {
    side(ic, 1, s0); side(ic + 1, 0, s1);
    if (P.riemann == 1) riemann_llf(s0, s1, F1, gamma); else riemann_hllc(s0, s1, F1, gamma);
  }
  for (int i = 0; i < n; ++i)
    for (int v = 0; v < NV; ++v) {
      double fv = 0.0, fve = 0.0, sv = 0.0, sve = 0.0;
      for (int j = 0; j < n; ++j) {
        fv = fv + fq[j][v] * B.dP[j][i] * B.wq[j];
        fve = fve + fqe[j][v] * B.dP[j][i] * B.wq[j];
        if (P.source == 2) {
          sv = sv + sq[j][v] * B.P[j][i] * B.wq[j];
          sve = sve + sqe[j][v] * B.P[j][i] * B.wq[j];
        }
      }
      M3(dudt, v, i, ic) = oneoverdx * fv - oneoverdx * fve - oneoverdx * (F1[v] * B.Ep[i] - F0[v] * B.Em[i])
                           + oneoverdx * (ffe[1][v] * B.Ep[i] - ffe[0][v] * B.Em[i]) + sv - sve;
    }
}
// limiter_TDV :520-600 with use_limiter = .false. on delta = w - q: w = q + limited(w - q)   (dg_with_source.f90:232-235)
__global__ void k_dg1_limiter_tdv(double* __restrict__ w, const double* __restrict__ q, double* __restrict__ delta_out, P1d P,
                                  const CtrlD* ctrl) {
  if (ctrl && ctrl->skip) return;
  int ic = blockIdx.x * blockDim.x + threadIdx.x;
  if (ic >= P.nx) return;
  const int n = P.n;
  double d[MAXN][NV];
  for (int i = 0; i < n; ++i)
    for (int v = 0; v < NV; ++v) d[i][v] = M3(w, v, i, ic) - M3(q, v, i, ic);
  if (n > 1) {
    double ul[NV] = {0, 0, 0}, ur[NV] = {0, 0, 0};
    for (int i = 1; i <= n; ++i)
      for (int v = 0; v < NV; ++v) {
        ul[v] = ul[v] + d[i - 1][v] * ((i - 1) % 2 == 0 ? 1.0 : -1.0) * sqrt(2.0 * (double)i - 1.0);
        ur[v] = ur[v] + d[i - 1][v] * sqrt(2.0 * (double)i - 1.0);
      }
    if (ul[0] < 1e-10 || ur[0] < 1e-10 || ul[2] < 1e-10 || ul[2] < 1e-10)
      for (int i = 1; i < n; ++i)
        for (int v = 0; v < NV; ++v) d[i][v] = 0.0;
  }
  for (int i = 0; i < n; ++i)
    for (int v = 0; v < NV; ++v) {
      M3(w, v, i, ic) = M3(q, v, i, ic) + d[i][v];
      if (delta_out) M3(delta_out, v, i, ic) = d[i][v];
    }
}
// limiter_cons :602-734: reads `u` (with neighbours), writes `ul`
__global__ void k_dg1_limiter_cons(const double* __restrict__ u, double* __restrict__ ul, P1d P, const CtrlD* ctrl) {
  if (ctrl && ctrl->skip) return;
  int ic = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (ic > P.nx) return;
  const int n = P.n, nx = P.nx;
  const double gamma = P.gamma;
  double el[MAXN][NV];
  for (int i = 0; i < n; ++i)
    for (int v = 0; v < NV; ++v) el[i][v] = M3(u, v, i, ic - 1);
  if (n > 1 && P.use_limiter) {
    int ileft = ic - 1, iright = ic + 1;
    double switch_left = 1.0, switch_right = 1.0;
    if (P.bc == 1) { if (ic == 1) ileft = nx; if (ic == nx) iright = 1; }
    if (P.bc == 2 || P.bc == 4) { if (ic == 1) ileft = 1; if (ic == nx) iright = nx; }
    if (P.bc == 3) { if (ic == 1) { ileft = 1; switch_left = -1.0; } if (ic == nx) { iright = nx; switch_right = -1.0; } }
    if (ileft >= 1 && iright <= nx) {
      double wL[MAXN][NV], wM[MAXN][NV], wR[MAXN][NV], w_lim[MAXN][NV];
      for (int i = n - 1; i >= 1; --i) {
        double coeff_i = sqrt(2.0 * (double)(i - 1) + 1.0) * (2.0 * (double)i - 1);
        double coeff_ip1 = sqrt(2.0 * (double)i + 1.0) * (2.0 * (double)i - 1);
        for (int v = 0; v < NV; ++v) {
          wL[i][v] = (M3(u, v, i - 1, ic - 1) - M3(u, v, i - 1, ileft - 1)) * coeff_i / coeff_ip1;
          wR[i][v] = (M3(u, v, i - 1, iright - 1) - M3(u, v, i - 1, ic - 1)) * coeff_i / coeff_ip1;
          wM[i][v] = M3(u, v, i, ic - 1);
        }
        wL[i][1] = switch_left * wL[i][1];
        wR[i][1] = switch_right * wR[i][1];
      }
      for (int i = 1; i < n; ++i)
        for (int v = 0; v < NV; ++v) w_lim[i][v] = wM[i][v];
      for (int v = 0; v < NV; ++v)
        for (int i = n - 1; i >= 1; --i) {
          double w_min = minmod3(wL[i][v], wM[i][v], wR[i][v]);
          w_lim[i][v] = w_min;
          if (fabs(w_min - wM[i][v]) < (double)0.01f * fabs(wM[i][v])) break;
        }
      for (int i = n - 1; i >= 1; --i)
        for (int v = 0; v < NV; ++v) el[i][v] = w_lim[i][v];
    }
  }
  if (n > 1) {
    double w[NV], u_left[NV] = {0, 0, 0}, u_right[NV] = {0, 0, 0}, w_left[NV], w_right[NV];
    prim(el[0], w, gamma);
    for (int i = 1; i <= n; ++i)
      for (int v = 0; v < NV; ++v) {
        u_left[v] = u_left[v] + el[i - 1][v] * ((i - 1) % 2 == 0 ? 1.0 : -1.0) * sqrt(2.0 * (double)i - 1.0);
        u_right[v] = u_right[v] + el[i - 1][v] * sqrt(2.0 * (double)i - 1.0);
      }
    cons_to_prim(u_left, w_left, w, gamma);
    cons_to_prim(u_right, w_right, w, gamma);
    if (w_left[0] < 1e-10 || w_right[0] < 1e-10 || w_left[2] < 1e-10 || w_left[2] < 1e-10)
      for (int i = 1; i < n; ++i)
        for (int v = 0; v < NV; ++v) el[i][v] = 0.0;
  }
  for (int i = 0; i < n; ++i)
    for (int v = 0; v < NV; ++v) M3(ul, v, i, ic - 1) = el[i][v];
}

// out = c0*A0 [+ c1*A1] [+ (cd*dt)*D], left to right (cd == 0: no D term at all)
__global__ void k_dg1_axpy(double* out, const double* A0, double c0, const double* A1, double c1, const double* D, double cd,
                           int n, int na, const CtrlD* ctrl) {
  if (ctrl->skip) return;
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double cdt = cd * ctrl->dt;
  double r = (c0 == 1.0) ? A0[k] : c0 * A0[k];
  if (na == 2) r = r + c1 * A1[k];
  out[k] = (cd == 0.0) ? r : r + cdt * D[k];
}
__global__ void k_dg1_advance(CtrlD* c) { if (c->skip) return; c->t = c->t + c->dt; c->iter = c->iter + 1; }
__global__ void k_dg1_ctrl_init(CtrlD* c, double tend, int max_iter) {
  c->t = 0.0; c->iter = 0; c->dt = 0.0; c->tend = tend; c->max_iter = max_iter; c->skip = 0;
}

}}  // namespace wb::dg1d

using namespace wb;
using namespace wb::dg1d;

struct wb_dg1d {
  P1d P;
  Basis1 B;
  int dev = 0;
  cudaStream_t stream = nullptr;
  size_t N = 0;
  double *du = nullptr, *ueq = nullptr, *uinit = nullptr, *dudt = nullptr, *w1 = nullptr, *w2 = nullptr, *w3 = nullptr, *w4 = nullptr, *w5 = nullptr, *w6 = nullptr;
  CtrlD* ctrl = nullptr;
  CtrlD* h_ctrl = nullptr;
};

#define F32(x) ((double)(x##f))

extern "C" {

int wb_dg1d_create(wb_dg1d** out, const wb_dg1d_params* p) {
  if (!out || !p) { set_error("null argument"); return WB_ERR_ARG; }
  *out = nullptr;
  WB_REQUIRE(p->nvar == 3, "nvar must be 3 (got %d)", p->nvar);
  WB_REQUIRE(p->n >= 1 && p->n <= 3, "n must be 1..3 (the hard-coded quadrature rules of the root legendre.f90)");
  WB_REQUIRE(p->nx >= 3, "nx must be >= 3");
  WB_REQUIRE(p->riemann == 1 || p->riemann == 2, "riemann must be 1 (llf) or 2 (hllc)");
  WB_REQUIRE(p->source == 1 || p->source == 2, "source must be 1 or 2");
  WB_REQUIRE(p->gamma > 1.0 && p->boxlen > 0, "gamma>1, boxlen>0 required");
  WB_REQUIRE(p->bc >= 1 && p->bc <= 5, "bc must be 1..5");
  int dev = 0;
  WB_CHECK(select_device(p->device, &dev));
  wb_dg1d* h = new wb_dg1d;
  h->dev = dev;
  h->P.n = p->n; h->P.nx = p->nx; h->P.riemann = p->riemann; h->P.source = p->source; h->P.gamma = p->gamma; h->P.boxlen = p->boxlen;
  h->P.bc = p->bc; h->P.use_limiter = p->use_limiter ? 1 : 0;
  h->B = make_basis(p->n);
  h->N = (size_t)NV * p->n * p->nx;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  double** bufs[] = {&h->du, &h->ueq, &h->uinit, &h->dudt, &h->w1, &h->w2, &h->w3, &h->w4, &h->w5, &h->w6};
  for (double** b : bufs)
    if (e == cudaSuccess) e = cudaMalloc(b, sizeof(double) * h->N);
  if (e == cudaSuccess) e = cudaMalloc(&h->ctrl, sizeof(CtrlD));
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_ctrl, sizeof(CtrlD));
  if (e == cudaSuccess) e = cudaMemset(h->ctrl, 0, sizeof(CtrlD));
  if (e != cudaSuccess) { set_error("dg1d allocation failed: %s", cudaGetErrorString(e)); wb_dg1d_destroy(h); return WB_ERR_CUDA; }
  *out = h;
  return WB_OK;
}
int wb_dg1d_destroy(wb_dg1d* h) {
  if (!h) return WB_OK;
  cudaSetDevice(h->dev);
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  cudaFree(h->du); cudaFree(h->ueq); cudaFree(h->uinit); cudaFree(h->dudt); cudaFree(h->w1); cudaFree(h->w2); cudaFree(h->w3);
  cudaFree(h->w4); cudaFree(h->w5); cudaFree(h->w6); cudaFree(h->ctrl);
  if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
  delete h;
  return WB_OK;
}
int wb_dg1d_quadrature(wb_dg1d* h, double* x, double* w) {
  if (!h || !x || !w) { set_error("null argument"); return WB_ERR_ARG; }
  for (int i = 0; i < h->P.n; ++i) { x[i] = h->B.xq[i]; w[i] = h->B.wq[i]; }
  return WB_OK;
}
int wb_dg1d_compute_update_exact_delta(wb_dg1d* h, const double* delta_u, const double* u_eq, double* dudt) {
  if (!h || !delta_u || !u_eq || !dudt) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  const size_t fb = sizeof(double) * h->N;
  WB_CUDA(cudaMemcpyAsync(h->du, delta_u, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CUDA(cudaMemcpyAsync(h->ueq, u_eq, fb, cudaMemcpyHostToDevice, h->stream));
  k_dg1_update<<<(h->P.nx + 63) / 64, 64, 0, h->stream>>>(h->du, h->ueq, h->dudt, h->P, h->B, nullptr);
  WB_LAUNCH_CHECK();
  WB_CUDA(cudaMemcpyAsync(dudt, h->dudt, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  return WB_OK;
}
int wb_dg1d_compute_max_speed(wb_dg1d* h, const double* u_nodes, double* cmax) {
  if (!h || !u_nodes || !cmax) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  WB_CUDA(cudaMemcpyAsync(h->uinit, u_nodes, sizeof(double) * h->N, cudaMemcpyHostToDevice, h->stream));
  k_dg1_max_speed<<<1, 256, 0, h->stream>>>(h->uinit, h->P, h->ctrl, 0);
  WB_LAUNCH_CHECK();
  WB_CUDA(cudaMemcpyAsync(h->h_ctrl, h->ctrl, sizeof(CtrlD), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  *cmax = h->h_ctrl->cmax;
  return WB_OK;
}
// main loop with integrator 'RKi' (:173-336, :282-305)
int wb_dg1d_evolve(wb_dg1d* h, double* delta_u, const double* u_eq, double* uinit, double tend, int max_iter, int* iters,
                   double* t_out, double* dt_out) {
  if (!h || !delta_u || !u_eq || !uinit) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  const size_t fb = sizeof(double) * h->N;
  const int N = (int)h->N;
  WB_CUDA(cudaMemcpyAsync(h->du, delta_u, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CUDA(cudaMemcpyAsync(h->ueq, u_eq, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CUDA(cudaMemcpyAsync(h->uinit, uinit, fb, cudaMemcpyHostToDevice, h->stream));
  k_dg1_ctrl_init<<<1, 1, 0, h->stream>>>(h->ctrl, tend, max_iter);
  WB_LAUNCH_CHECK();
  dim3 bc(64), gc((h->P.nx + 63) / 64), ba(128), ga((N + 127) / 128);
  auto upd = [&](const double* in) {
    k_dg1_update<<<gc, bc, 0, h->stream>>>(in, h->ueq, h->dudt, h->P, h->B, h->ctrl);
    wb::g_launches.fetch_add(1);
  };
  auto axpy = [&](double* out, const double* A0, double c0, const double* A1, double c1, double cd, int na) {
    k_dg1_axpy<<<ga, ba, 0, h->stream>>>(out, A0, c0, A1, c1, h->dudt, cd, N, na, h->ctrl);
    wb::g_launches.fetch_add(1);
  };
  int it = 0;
  double t = 0.0, dt = 0.0;
  for (;;) {
    if (!(t < tend) || (max_iter >= 0 && it >= max_iter)) break;
    for (int s = 0; s < 16; ++s) {
      k_dg1_max_speed<<<1, 256, 0, h->stream>>>(h->uinit, h->P, h->ctrl, 1);
      wb::g_launches.fetch_add(1);
      upd(h->du);
      axpy(h->w1, h->du, 1.0, nullptr, 0.0, F32(0.391752226571890), 1);
      upd(h->w1);
      axpy(h->w2, h->du, F32(0.444370493651235), h->w1, F32(0.555629506348765), F32(0.368410593050371), 2);
      upd(h->w2);
      axpy(h->w3, h->du, F32(0.620101851488403), h->w2, F32(0.379898148511597), F32(0.251891774271694), 2);
      upd(h->w3);
      axpy(h->w4, h->du, F32(0.178079954393132), h->w3, F32(0.821920045606868), F32(0.544974750228521), 2);
      axpy(h->du, h->w2, F32(0.517231671970585), h->w3, F32(0.096059710526147), F32(0.063692468666290), 2);
      upd(h->w4);
      axpy(h->du, h->du, 1.0, h->w4, F32(0.386708617503269), F32(0.226007483236906), 2);
      k_dg1_reconstruct<<<gc, bc, 0, h->stream>>>(h->du, h->ueq, h->uinit, h->P, h->B, h->ctrl);
      k_dg1_advance<<<1, 1, 0, h->stream>>>(h->ctrl);
      wb::g_launches.fetch_add(2);
    }
    WB_CUDA(cudaGetLastError());
    WB_CUDA(cudaMemcpyAsync(h->h_ctrl, h->ctrl, sizeof(CtrlD), cudaMemcpyDeviceToHost, h->stream));
    WB_CUDA(cudaStreamSynchronize(h->stream));
    it = h->h_ctrl->iter; t = h->h_ctrl->t; dt = h->h_ctrl->dt;
  }
  WB_CUDA(cudaMemcpyAsync(delta_u, h->du, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaMemcpyAsync(uinit, h->uinit, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (iters) *iters = it;
  if (t_out) *t_out = t;
  if (dt_out) *dt_out = dt;
  return WB_OK;
}

// replaces compute_update(u,dudt)   dg_with_source.f90:807-1028
int wb_dg1d_compute_update(wb_dg1d* h, const double* u, double* dudt) {
  if (!h || !u || !dudt) { set_error("null argument"); return WB_ERR_ARG; }
  WB_REQUIRE(h->P.bc >= 1 && h->P.bc <= 4, "compute_update needs bc 1..4: with bc = 5 the reference reads u_right(:,0) / u_left(:,nx+1) out of bounds");
  WB_CUDA(cudaSetDevice(h->dev));
  const size_t fb = sizeof(double) * h->N;
  WB_CUDA(cudaMemcpyAsync(h->du, u, fb, cudaMemcpyHostToDevice, h->stream));
  k_dg1_update_plain<<<(h->P.nx + 63) / 64, 64, 0, h->stream>>>(h->du, h->dudt, h->P, h->B, nullptr);
  WB_LAUNCH_CHECK();
  WB_CUDA(cudaMemcpyAsync(dudt, h->dudt, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  return WB_OK;
}
// replaces limiter(u)   dg_with_source.f90:414-519
int wb_dg1d_limiter(wb_dg1d* h, double* u_inout) {
  if (!h || !u_inout) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  const size_t fb = sizeof(double) * h->N;
  WB_CUDA(cudaMemcpyAsync(h->du, u_inout, fb, cudaMemcpyHostToDevice, h->stream));
  if (h->P.n > 1) {
    k_dg1_limiter<<<(h->P.nx + 63) / 64, 64, 0, h->stream>>>(h->du, h->w1, h->P, nullptr);
    WB_LAUNCH_CHECK();
    WB_CUDA(cudaMemcpyAsync(u_inout, h->w1, fb, cudaMemcpyDeviceToHost, h->stream));
  }
  WB_CUDA(cudaStreamSynchronize(h->stream));
  return WB_OK;
}
// main loop with integrator 'RK1' (1), 'RK2' (2), 'RK3' (3), 'RK4' (4)   dg_with_source.f90:173-227, :313-336
int wb_dg1d_evolve_rk(wb_dg1d* h, int integrator, double* u, const double* delta_u, const double* u_eq, double* uinit, double tend,
                      int max_iter, int* iters, double* t_out, double* dt_out) {
  if (!h || !u || !delta_u || !u_eq || !uinit) { set_error("null argument"); return WB_ERR_ARG; }
  WB_REQUIRE(integrator >= 1 && integrator <= 4, "integrator must be 1..4 ('RK1'..'RK4'); 'RKi' is wb_dg1d_evolve");
  WB_REQUIRE(h->P.bc >= 1 && h->P.bc <= 4, "the plain update needs bc 1..4");
  WB_CUDA(cudaSetDevice(h->dev));
  const size_t fb = sizeof(double) * h->N;
  const int N = (int)h->N;
  // buffers: w5 holds u, du holds delta_u (constant on these paths), w6 is the limiter scratch
  double *U = h->w5, *W1 = h->w1, *W2 = h->w2, *W3 = h->w3, *W4 = h->w4, *S = h->w6;
  WB_CUDA(cudaMemcpyAsync(U, u, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CUDA(cudaMemcpyAsync(h->du, delta_u, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CUDA(cudaMemcpyAsync(h->ueq, u_eq, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CUDA(cudaMemcpyAsync(h->uinit, uinit, fb, cudaMemcpyHostToDevice, h->stream));
  k_dg1_ctrl_init<<<1, 1, 0, h->stream>>>(h->ctrl, tend, max_iter);
  WB_LAUNCH_CHECK();
  dim3 bc(64), gc((h->P.nx + 63) / 64), ba(128), ga((N + 127) / 128);
  auto upd = [&](const double* in) {
    k_dg1_update_plain<<<gc, bc, 0, h->stream>>>(in, h->dudt, h->P, h->B, h->ctrl);
    wb::g_launches.fetch_add(1);
  };
  auto axpy = [&](double* out, const double* A0, double c0, const double* A1, double c1, double cd, int na) {
    k_dg1_axpy<<<ga, ba, 0, h->stream>>>(out, A0, c0, A1, c1, h->dudt, cd, N, na, h->ctrl);
    wb::g_launches.fetch_add(1);
  };
  auto lim = [&](double* x) {       // limiter reads neighbours: limit into the scratch, then copy back (skipped steps copy nothing)
    if (h->P.n == 1) return;
    k_dg1_limiter<<<gc, bc, 0, h->stream>>>(x, S, h->P, h->ctrl);
    k_dg1_axpy<<<ga, ba, 0, h->stream>>>(x, S, 1.0, nullptr, 0.0, S, 0.0, N, 1, h->ctrl);
    wb::g_launches.fetch_add(2);
  };
  int it = 0;
  double t = 0.0, dt = 0.0;
  for (;;) {
    if (!(t < tend) || (max_iter >= 0 && it >= max_iter)) break;
    for (int s = 0; s < 16; ++s) {
      k_dg1_max_speed<<<1, 256, 0, h->stream>>>(h->uinit, h->P, h->ctrl, 1);
      wb::g_launches.fetch_add(1);
      if (integrator == 1) {
        upd(U); axpy(U, U, 1.0, nullptr, 0.0, 1.0, 1);
      } else if (integrator == 2) {
        upd(U); axpy(W1, U, 1.0, nullptr, 0.0, 1.0, 1); lim(W1);
        upd(W1); axpy(U, U, 0.5, W1, 0.5, 0.5, 2); lim(U);
      } else if (integrator == 3) {
        upd(U); axpy(W1, U, 1.0, nullptr, 0.0, 1.0, 1); lim(W1);
        upd(W1); axpy(W2, U, 0.75, W1, 0.25, 0.25, 2); lim(W2);
        upd(W2); axpy(U, U, (double)(1.0f / 3.0f), W2, (double)(2.0f / 3.0f), (double)(2.0f / 3.0f), 2); lim(U);
      } else {
        // :206 u = u - u_eq (the MODES minus the NODAL equilibrium, as shipped) ... :226 u = u + u_eq
        axpy(U, U, 1.0, h->ueq, -1.0, 0.0, 2);
        upd(U); axpy(W1, U, 1.0, nullptr, 0.0, F32(0.391752226571890), 1); lim(W1);
        upd(W1); axpy(W2, U, F32(0.444370493651235), W1, F32(0.555629506348765), F32(0.368410593050371), 2); lim(W2);
        upd(W2); axpy(W3, U, F32(0.620101851488403), W2, F32(0.379898148511597), F32(0.251891774271694), 2); lim(W3);
        upd(W3); axpy(W4, U, F32(0.178079954393132), W3, F32(0.821920045606868), F32(0.544974750228521), 2);
        axpy(U, W2, F32(0.517231671970585), W3, F32(0.096059710526147), F32(0.063692468666290), 2); lim(W4);
        upd(W4); axpy(U, U, 1.0, W4, F32(0.386708617503269), F32(0.226007483236906), 2); lim(U);
        axpy(U, U, 1.0, h->ueq, 1.0, 0.0, 2);
      }
      k_dg1_reconstruct<<<gc, bc, 0, h->stream>>>(h->du, h->ueq, h->uinit, h->P, h->B, h->ctrl);
      k_dg1_advance<<<1, 1, 0, h->stream>>>(h->ctrl);
      wb::g_launches.fetch_add(2);
    }
    WB_CUDA(cudaGetLastError());
    WB_CUDA(cudaMemcpyAsync(h->h_ctrl, h->ctrl, sizeof(CtrlD), cudaMemcpyDeviceToHost, h->stream));
    WB_CUDA(cudaStreamSynchronize(h->stream));
    it = h->h_ctrl->iter; t = h->h_ctrl->t; dt = h->h_ctrl->dt;
  }
  WB_CUDA(cudaMemcpyAsync(u, U, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaMemcpyAsync(uinit, h->uinit, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (iters) *iters = it;
  if (t_out) *t_out = t;
  if (dt_out) *dt_out = dt;
  return WB_OK;
}


// replaces compute_update_exact(u,u_eq_modes,dudt)   dg_with_source.f90:1380-1744 (bc 4 | 5)
int wb_dg1d_compute_update_exact(wb_dg1d* h, const double* u, const double* u_eq_modes, double* dudt) {
  if (!h || !u || !u_eq_modes || !dudt) { set_error("null argument"); return WB_ERR_ARG; }
  WB_REQUIRE(h->P.bc == 4 || h->P.bc == 5, "compute_update_exact is defined for bc 4 and 5 only (other bc use out-of-bounds reads in the reference)");
  WB_CUDA(cudaSetDevice(h->dev));
  const size_t fb = sizeof(double) * h->N;
  WB_CUDA(cudaMemcpyAsync(h->du, u, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CUDA(cudaMemcpyAsync(h->ueq, u_eq_modes, fb, cudaMemcpyHostToDevice, h->stream));
  k_dg1_update_exact<<<(h->P.nx + 63) / 64, 64, 0, h->stream>>>(h->du, h->ueq, h->dudt, h->P, h->B, nullptr);
  WB_LAUNCH_CHECK();
  WB_CUDA(cudaMemcpyAsync(dudt, h->dudt, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  return WB_OK;
}
// replaces limiter_TDV(u) (:520-600, use_limiter = .false.: positivity fallback on the traces) and limiter_cons(u) (:602-734)
int wb_dg1d_limiter_tdv(wb_dg1d* h, double* u_inout) {
  if (!h || !u_inout) { set_error("null argument"); return WB_ERR_ARG; }
  WB_REQUIRE(!h->P.use_limiter, "limiter_TDV with use_limiter = .true. indexes its neighbours with a stale loop variable in the "
                                "reference (dg_with_source.f90:550-552): undefined, not built");
  WB_CUDA(cudaSetDevice(h->dev));
  const size_t fb = sizeof(double) * h->N;
  WB_CUDA(cudaMemcpyAsync(h->w1, u_inout, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CUDA(cudaMemsetAsync(h->w2, 0, fb, h->stream));
  k_dg1_limiter_tdv<<<(h->P.nx + 63) / 64, 64, 0, h->stream>>>(h->w1, h->w2, nullptr, h->P, nullptr);
  WB_LAUNCH_CHECK();
  WB_CUDA(cudaMemcpyAsync(u_inout, h->w1, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  return WB_OK;
}
int wb_dg1d_limiter_cons(wb_dg1d* h, double* u_inout) {
  if (!h || !u_inout) { set_error("null argument"); return WB_ERR_ARG; }
  WB_CUDA(cudaSetDevice(h->dev));
  const size_t fb = sizeof(double) * h->N;
  WB_CUDA(cudaMemcpyAsync(h->du, u_inout, fb, cudaMemcpyHostToDevice, h->stream));
  k_dg1_limiter_cons<<<(h->P.nx + 63) / 64, 64, 0, h->stream>>>(h->du, h->w1, h->P, nullptr);
  WB_LAUNCH_CHECK();
  WB_CUDA(cudaMemcpyAsync(u_inout, h->w1, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  return WB_OK;
}
// main loop with integrator 'RKw' (5, :229-270) or 'RKe' (6, :273-280)   dg_with_source.f90:173-336
int wb_dg1d_evolve_w(wb_dg1d* h, int integrator, double* u, double* delta_u, const double* u_eq_nodes, const double* u_eq_modes,
                     double* uinit, double tend, int max_iter, int* iters, double* t_out, double* dt_out) {
  if (!h || !u || !delta_u || !u_eq_nodes || !u_eq_modes || !uinit) { set_error("null argument"); return WB_ERR_ARG; }
  WB_REQUIRE(integrator == 5 || integrator == 6, "integrator must be 5 ('RKw') or 6 ('RKe')");
  WB_REQUIRE(integrator == 6 || h->P.bc == 4 || h->P.bc == 5, "'RKw' (compute_update_exact) is defined for bc 4 and 5 only");
  WB_REQUIRE(integrator == 6 || !h->P.use_limiter, "'RKw' with use_limiter = .true. is undefined in the reference (limiter_TDV)");
  WB_CUDA(cudaSetDevice(h->dev));
  const size_t fb = sizeof(double) * h->N;
  const int N = (int)h->N;
  double *U = h->w5, *W1 = h->w1, *W2 = h->w2, *W3 = h->w3, *W4 = h->w4, *Q = h->w6, *D = h->du;
  WB_CUDA(cudaMemcpyAsync(U, u, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CUDA(cudaMemcpyAsync(D, delta_u, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CUDA(cudaMemcpyAsync(h->ueq, u_eq_nodes, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CUDA(cudaMemcpyAsync(Q, u_eq_modes, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CUDA(cudaMemcpyAsync(h->uinit, uinit, fb, cudaMemcpyHostToDevice, h->stream));
  k_dg1_ctrl_init<<<1, 1, 0, h->stream>>>(h->ctrl, tend, max_iter);
  WB_LAUNCH_CHECK();
  dim3 bc(64), gc((h->P.nx + 63) / 64), ba(128), ga((N + 127) / 128);
  auto updw = [&](const double* in) {
    k_dg1_update_exact<<<gc, bc, 0, h->stream>>>(in, Q, h->dudt, h->P, h->B, h->ctrl);
    wb::g_launches.fetch_add(1);
  };
  auto updd = [&](const double* in) {
    k_dg1_update<<<gc, bc, 0, h->stream>>>(in, h->ueq, h->dudt, h->P, h->B, h->ctrl);
    wb::g_launches.fetch_add(1);
  };
  auto axpy = [&](double* out, const double* A0, double c0, const double* A1, double c1, double cd, int na) {
    k_dg1_axpy<<<ga, ba, 0, h->stream>>>(out, A0, c0, A1, c1, h->dudt, cd, N, na, h->ctrl);
    wb::g_launches.fetch_add(1);
  };
  auto limw = [&](double* x, double* delta_out) {     // delta = x - q; limiter_TDV(delta); x = q + delta
    k_dg1_limiter_tdv<<<gc, bc, 0, h->stream>>>(x, Q, delta_out, h->P, h->ctrl);
    wb::g_launches.fetch_add(1);
  };
  auto limc = [&](double* x, double* scratch) {       // limiter_cons reads neighbours: limit into the scratch, copy back
    if (h->P.n == 1) return;
    k_dg1_limiter_cons<<<gc, bc, 0, h->stream>>>(x, scratch, h->P, h->ctrl);
    k_dg1_axpy<<<ga, ba, 0, h->stream>>>(x, scratch, 1.0, nullptr, 0.0, scratch, 0.0, N, 1, h->ctrl);
    wb::g_launches.fetch_add(2);
  };
  int it = 0;
  double t = 0.0, dt = 0.0;
  for (;;) {
    if (!(t < tend) || (max_iter >= 0 && it >= max_iter)) break;
    for (int s = 0; s < 16; ++s) {
      k_dg1_max_speed<<<1, 256, 0, h->stream>>>(h->uinit, h->P, h->ctrl, 1);
      wb::g_launches.fetch_add(1);
      if (integrator == 5) {
        updw(U); axpy(W1, U, 1.0, nullptr, 0.0, F32(0.391752226571890), 1); limw(W1, nullptr);
        updw(W1); axpy(W2, U, F32(0.444370493651235), W1, F32(0.555629506348765), F32(0.368410593050371), 2); limw(W2, nullptr);
        updw(W2); axpy(W3, U, F32(0.620101851488403), W2, F32(0.379898148511597), F32(0.251891774271694), 2); limw(W3, nullptr);
        updw(W3); axpy(W4, U, F32(0.178079954393132), W3, F32(0.821920045606868), F32(0.544974750228521), 2);
        axpy(U, W2, F32(0.517231671970585), W3, F32(0.096059710526147), F32(0.063692468666290), 2); limw(W4, nullptr);
        updw(W4); axpy(U, U, 1.0, W4, F32(0.386708617503269), F32(0.226007483236906), 2); limw(U, D);
      } else {
        updd(D); limc(D, W2); axpy(W1, D, 1.0, nullptr, 0.0, 1.0, 1);
        updd(W1); limc(W1, W2); axpy(D, D, 0.5, W1, 0.5, 0.5, 2);
      }
      k_dg1_reconstruct<<<gc, bc, 0, h->stream>>>(D, h->ueq, h->uinit, h->P, h->B, h->ctrl);
      k_dg1_advance<<<1, 1, 0, h->stream>>>(h->ctrl);
      wb::g_launches.fetch_add(2);
    }
    WB_CUDA(cudaGetLastError());
    WB_CUDA(cudaMemcpyAsync(h->h_ctrl, h->ctrl, sizeof(CtrlD), cudaMemcpyDeviceToHost, h->stream));
    WB_CUDA(cudaStreamSynchronize(h->stream));
    it = h->h_ctrl->iter; t = h->h_ctrl->t; dt = h->h_ctrl->dt;
  }
  WB_CUDA(cudaMemcpyAsync(u, U, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaMemcpyAsync(delta_u, D, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaMemcpyAsync(uinit, h->uinit, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (iters) *iters = it;
  if (t_out) *t_out = t;
  if (dt_out) *dt_out = dt;
  return WB_OK;
}

}  // extern "C"
