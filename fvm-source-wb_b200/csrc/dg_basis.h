// Host-side Legendre basis and quadrature tables of the 2D DG path, computed with the reference's own
// recurrences (2d/legendre.f90) -- Newton iteration for the Gauss-Legendre nodes, tabulated GLL points as
// shipped (including the odd n = 2, 3 weights) -- and handed to the kernels as a by-value struct.
#pragma once
#include <cmath>
#include <cstring>

namespace wb { namespace dg {

constexpr int MAXM = 4;   // supported order per direction (mx = my <= 4)

// P_n(x)*sqrt(2n+1); the Fortran function clamps its argument to [-1,1] in place (2d/legendre.f90:5)
inline double legendre(double& x, int n) {
  x = std::fmin(std::fmax(x, (double)-1.0f), (double)1.0f);
  double l = 0.0;
  switch (n) {
    case 0: l = 1.0; break;
    case 1: l = x; break;
    case 2: l = 0.5 * (3 * (x * x) - 1); break;
    case 3: l = 0.5 * (5.0 * ((x * x) * x) - 3.0 * x); break;
    case 4: { double x2 = x * x; l = 0.125 * (35.0 * (x2 * x2) - 30.0 * x2 + 3.0); } break;
    default: break;
  }
  return std::sqrt((2.0 * (double)n + 1.0)) * l;
}
// 2d/legendre.f90:27-50
inline double legendre_prime(double& x, int n) {
  x = std::fmin(std::fmax(x, (double)-1.0f), (double)1.0f);
  double l = 0.0;
  switch (n) {
    case 0: l = 0.0; break;
    case 1: l = 1.0; break;
    case 2: l = 3.0 * x; break;
    case 3: l = 0.5 * (15.0 * (x * x) - 3.0); break;
    case 4: l = 0.125 * (140.0 * ((x * x) * x) - 60.0 * x); break;
    default: break;
  }
  return std::sqrt((2.0 * (double)n + 1.0)) * l;
}
// 2d/legendre.f90:77-108: Newton (500 iterations) from a single-precision initial factor, then mirrored
inline void gl_quadrature(double* x_quad, double* w_quad, int n) {
  const double dpi = std::acos(-1.0);
  for (int i = 1; i <= n; ++i) {
    float fn = (float)n;
    float pre = (1.0f - 0.125f / fn / fn) + 0.125f / fn / fn / fn;
    double xx = (double)pre * std::cos(dpi * (4.0 * (double)i - 1.0) / (4.0 * (double)n + 2.0));
    for (int iter = 1; iter <= 500; ++iter) {
      double a = legendre(xx, n);
      double b = legendre_prime(xx, n);
      xx = xx - a / b;
    }
    double xi = -xx;
    double lp = legendre_prime(xi, n);
    x_quad[i - 1] = xi;
    w_quad[i - 1] = 2 * (2.0 * (double)n + 1.0) / (1.0 - xi * xi) / (lp * lp);
  }
  for (int i = n / 2 + 1; i <= n; ++i) {
    x_quad[i - 1] = -x_quad[n - i];
    w_quad[i - 1] = w_quad[n - i];
  }
}
// 2d/legendre.f90:111-170 (real(4) arithmetic, values as shipped)
inline void gll_quadrature(double* x, double* w, int n) {
  switch (n) {
    case 2: x[0] = -1.; w[0] = 1.; x[1] = 0.; w[1] = 1.; break;
    case 3: x[0] = -1.; w[0] = 3.f / 4.f; x[2] = 1.; w[2] = 3.f / 4.f; x[1] = 0.; w[1] = 1.f / 4.f; break;
    case 4:
      x[0] = -1.; w[0] = 1.f / 6.f; x[1] = -1.f / 5.f * sqrtf(5.f); w[1] = 5.f / 6.f;
      x[2] = 1.f / 5.f * sqrtf(5.f); w[2] = 5.f / 6.f; x[3] = 1.; w[3] = 1.f / 6.f; break;
    default: break;
  }
}

struct Basis {
  int m, gll;
  double xq[MAXM], wq[MAXM];      // gl_quadrature(x_quad, w_x_quad, mx) (== the y rule since mx == my)
  double P[MAXM][MAXM];           // P[q][n]  = legendre(x_quad(q), n)
  double dP[MAXM][MAXM];          // dP[q][n] = legendre_prime(x_quad(q), n)
  double Em[MAXM], Ep[MAXM];      // legendre(-1, n), legendre(+1, n)
  double xg[MAXM], Pg[MAXM][MAXM];// GLL points and legendre(x_gll(r), n)
};

inline Basis make_basis(int m) {
  Basis B;
  std::memset(&B, 0, sizeof(B));
  B.m = m;
  B.gll = (2 * (m - 1) + 3) / 2;
  gl_quadrature(B.xq, B.wq, m);
  double wg[MAXM] = {0, 0, 0, 0};
  if (B.gll >= 2) gll_quadrature(B.xg, wg, B.gll);
  for (int q = 0; q < m; ++q)
    for (int n = 0; n < m; ++n) {
      double x = B.xq[q];
      B.P[q][n] = legendre(x, n);
      x = B.xq[q];
      B.dP[q][n] = legendre_prime(x, n);
    }
  for (int n = 0; n < m; ++n) {
    double a = -1.0, b = 1.0;
    B.Em[n] = legendre(a, n);
    B.Ep[n] = legendre(b, n);
  }
  for (int r = 0; r < B.gll; ++r)
    for (int n = 0; n < m; ++n) { double x = B.xg[r]; B.Pg[r][n] = legendre(x, n); }
  return B;
}

}}  // namespace wb::dg
