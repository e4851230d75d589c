/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement (plain C, FP64, -ffp-contract=off) of the default path of the reference's 1D modal DG code
 * dg_with_source.f90 (module dg_commons.f90, basis in the ROOT legendre.f90): integrator 'RKi' = SSPRK(5,4) on the
 * perturbation delta_u with compute_update_exact_delta (:1749-2031), riemann_hllc (:1318-1374, default riemann=2) or
 * riemann_llf (:1299-1316), source term, nodal reconstruction and time-step control of the main loop (:173-336).
 * Also restated: 'RK1'..'RK4' (compute_update + limiter), 'RKw' (compute_update_exact + limiter_TDV), 'RKe' (limiter_cons).
 *
 * PARITY PINNED TO THE REFERENCE'S OWN SOURCE TEXT: dg_with_source.f90 + the root legendre.f90 are EXECUTED, unmodified,
 * by the Fortran-90 interpreter oracle/f90interp.py (no Fortran compiler in the image); this file reproduces the vectors
 * (tests/golden/ref_dg1d.npz, generator tests/golden/make_ref_golden.py) BIT FOR BIT: the set-up of program dg, all
 * three update routines, the three limiters and the main loop with every integrator (tests/test_reference_pins.py).
 * Additional pins: the invariants of SURVEY section 4.2 (tests/test_oracle_dg1d.py).
 *
 * Layout: Fortran u(nvar,n,nx) == C double[nx][n][3].
 * Literal kinds: root legendre.f90 normalises P0..P2 with SINGLE-precision sqrt constants; its gl_quadrature
 * hard-codes n = 1,2,3 in single precision (:86-108); 0.9, 1.4, 1e-8 and the SSPRK(5,4) coefficients are real(4).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NV 3
#define MAXN 4
#define F32(x) ((double)(x##f))

typedef struct {
  int n;         /* dg_commons.f90:4 (nquad = n) */
  int nx;        /* :6 */
  int riemann;   /* :9   1 llf, 2 hllc */
  int source;    /* :15  1 none, 2 gravity */
  int ninit;     /* :13 */
  double gamma;  /* :18 */
  double boxlen; /* :17 */
  double pert;   /* :19 */
  int bc;          /* :14  1 periodic, 2 zero gradient, 3 reflective, 4 reflective on the end states; 5 (default) = none */
  int use_limiter; /* :10 */
} orc_dg1d_params;

/* root legendre.f90:1-25 (clamps its argument in place) */
double orc_dg1d_legendre(double *px, int n) {
  double x = fmin(fmax(*px, (double)-1.0f), (double)1.0f);
  *px = x;
  switch (n) {
    case 0: return (double)(1.0f * sqrtf(0.5f));
    case 1: return x * 0.5 * (double)sqrtf(6.f);
    case 2: return 0.25 * (3.0 * (x * x) - 1.0) * (double)sqrtf(10.f);
    case 3: return 0.5 * (5.0 * ((x * x) * x) - 3.0 * x);
    case 4: { double x2 = x * x; return 0.125 * (35.0 * (x2 * x2) - 30.0 * x2 + 3.0); }
    default: return 0.0;
  }
}
/* root legendre.f90:27-50 */
double orc_dg1d_legendre_prime(double *px, int n) {
  double x = fmin(fmax(*px, (double)-1.0f), (double)1.0f);
  *px = x;
  switch (n) {
    case 0: return 0.0;
    case 1: return (double)(1.0f * 0.5f * sqrtf(6.f));
    case 2: return 6.0 * x * 0.25 * (double)sqrtf(10.f);
    case 3: return 0.5 * (15.0 * (x * x) - 3.0);
    case 4: return 0.125 * (140.0 * ((x * x) * x) - 60.0 * x);
    default: return 0.0;
  }
}
static double leg(double x, int n) { return orc_dg1d_legendre(&x, n); }
static double legp(double x, int n) { return orc_dg1d_legendre_prime(&x, n); }

/* root legendre.f90:77-128 */
void orc_dg1d_gl_quadrature(double *x, double *w, int n) {
  if (n == 1) { x[0] = 0.0; w[0] = 2.0; return; }
  if (n == 2) {
    x[0] = -1.f / (double)3 * (double)sqrtf(3.f);
    x[1] = 1.f / (double)3 * (double)sqrtf(3.f);
    w[0] = 1.; w[1] = 1.;
    return;
  }
  if (n == 3) {
    x[0] = (double)(-sqrtf(3.f) / sqrtf(5.f)); x[1] = 0.0; x[2] = (double)(sqrtf(3.f) / sqrtf(5.f));
    w[0] = (double)(5.f / 9.f); w[1] = (double)(8.f / 9.f); w[2] = (double)(5.f / 9.f);
    return;
  }
  const double dpi = acos(-1.0);
  for (int i = 1; i <= n; ++i) {
    float fn = (float)n;
    float pre = (1.0f - 0.125f / fn / fn) + 0.125f / fn / fn / fn;
    double xx = (double)pre * cos(dpi * (4.0 * (double)i - 1.0) / (4.0 * (double)n + 2.0));
    for (int it = 1; it <= 50; ++it) { double a = orc_dg1d_legendre(&xx, n), b = orc_dg1d_legendre_prime(&xx, n); xx = xx - a / b; }
    x[i - 1] = xx;
    double lp = orc_dg1d_legendre_prime(&x[i - 1], n);
    w[i - 1] = 2.0 * (2.0 * (double)n + 1.0) / (1.0 - x[i - 1] * x[i - 1]) / (lp * lp);
  }
  for (int i = n / 2 + 1; i <= n; ++i) { x[i - 1] = -x[n - i]; w[i - 1] = w[n - i]; }
}

typedef struct { double xq[MAXN], wq[MAXN], P[MAXN][MAXN], dP[MAXN][MAXN], Em[MAXN], Ep[MAXN]; } basis1_t;
static void make_basis1(int n, basis1_t *B) {
  memset(B, 0, sizeof(*B));
  orc_dg1d_gl_quadrature(B->xq, B->wq, n);
  for (int q = 0; q < n; ++q)
    for (int m = 0; m < n; ++m) { B->P[q][m] = leg(B->xq[q], m); B->dP[q][m] = legp(B->xq[q], m); }
  for (int m = 0; m < n; ++m) { B->Em[m] = leg(-1.0, m); B->Ep[m] = leg(1.0, m); }
}
void orc_dg1d_quadrature(const orc_dg1d_params *p, double *x, double *w) { orc_dg1d_gl_quadrature(x, w, p->n); }

/* :1213-1223 / :1236-1246 / :1181-1194 / :1196-1211 / :1167-1179 */
static void prim(const double *u, double *w, double gamma) {
  w[0] = u[0];
  w[1] = u[1] / w[0];
  w[2] = (gamma - (double)1.0f) * (u[2] - 0.5 * w[0] * (w[1] * w[1]));
}
static void cons(const double *w, double *u, double gamma) {
  u[0] = w[0];
  u[1] = w[0] * w[1];
  u[2] = w[2] / (gamma - (double)1.0f) + 0.5 * w[0] * (w[1] * w[1]);
}
static void flux(const double *u, double *f, double gamma) {
  double w[NV];
  prim(u, w, gamma);
  f[0] = w[1] * u[0];
  f[1] = w[1] * u[1] + w[2];
  f[2] = w[1] * u[2] + w[2] * w[1];
}
static void source_term(const double *u, double *s, double gamma) {
  double w[NV];
  prim(u, w, gamma);
  s[0] = 0;
  s[1] = -w[0];
  s[2] = -w[0] * w[1];
}
static double speed(const double *u, double gamma) {
  double w[NV];
  prim(u, w, gamma);
  double cs = sqrt(gamma * fmax(w[2], 1e-10) / fmax(w[0], 1e-10));
  return fabs(w[1]) + cs;
}
/* :1299-1316 */
static void riemann_llf(const double *ul, const double *ur, double *fg, double gamma) {
  double cl = speed(ul, gamma), cr = speed(ur, gamma), cmax = fmax(cl, cr), fl[NV], fr[NV];
  flux(ul, fl, gamma);
  flux(ur, fr, gamma);
  for (int v = 0; v < NV; ++v) fg[v] = 0.5 * (fr[v] + fl[v]) - 0.5 * cmax * (ur[v] - ul[v]);
}
/* :1318-1374 */
static void riemann_hllc(const double *ul, const double *ur, double *fg, double gamma) {
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

/* :1053-1134 condinit */
void orc_dg1d_condinit(const orc_dg1d_params *p, double x, double *uu) {
  const double dpi = acos(-1.0), gamma = p->gamma;
  double ww[NV] = {0, 0, 0};
  switch (p->ninit) {
    case 1: ww[0] = 1.0 + 0.5 * sin(2.0 * dpi * x); ww[1] = 1.0; ww[2] = 1.0; break;
    case 2: ww[0] = (fabs(x - 0.5) < 0.25) ? 2. : 1.0; ww[1] = 1.0; ww[2] = 1.0; break;
    case 3:
      ww[0] = 1. + exp(-((x - 0.25) * (x - 0.25)) / 2.0 / (double)(0.05f * 0.05f)  /* 0.05**2 is real(4)**integer: folded in single precision */);
      if (fabs(x - (double)0.7f) < (double)0.1f) ww[0] = ww[0] + 1.;
      ww[1] = 1.0; ww[2] = 1.0; break;
    case 4:
      if (fabs(x - 0.25) < 0.25) { ww[0] = 1.0; ww[1] = 0.0; ww[2] = 1.0; }
      else { ww[0] = 0.125; ww[1] = 0.0; ww[2] = (double)0.1f; }
      break;
    case 5:
      if (x < (double)0.1f) { ww[0] = 1.0; ww[1] = 0.0; ww[2] = 1000.0; }
      else if (x < (double)0.9f) { ww[0] = 1.0; ww[1] = 0.0; ww[2] = (double)0.01f; }
      else { ww[0] = 1.0; ww[1] = 0.0; ww[2] = 100.; }
      break;
    case 6:
      if (x < 10.0) { ww[0] = (double)3.857143f; ww[1] = (double)-0.920279f; ww[2] = (double)10.333333f; }
      else { ww[0] = 1.0 + (double)0.2f * sin(5.0 * (x - 10.0)); ww[1] = (double)-3.549648f; ww[2] = 1.0; }
      break;
    case 7: ww[0] = exp(-x); ww[1] = 0.; ww[2] = exp(-x); break;
    default: { /* 8 */
      double d = x - p->boxlen / 2.;
      ww[0] = exp(-x); ww[1] = 0.; ww[2] = exp(-x) + p->pert * exp(-100 * (d * d));
    } break;
  }
  cons(ww, uu, gamma);
}

/* program dg :33-171: nodal IC `uinit`, nodal equilibrium `u_eq`, projected perturbation `delta_u` */
void orc_dg1d_setup(const orc_dg1d_params *p, double *uinit, double *u_eq, double *delta_u) {
  const int n = p->n, nx = p->nx;
  basis1_t B; make_basis1(n, &B);
  const double dx = p->boxlen / (double)nx;
  memset(delta_u, 0, sizeof(double) * NV * n * nx);
  for (int ic = 1; ic <= nx; ++ic) {
    double xcell = ((double)ic - 0.5) * dx;
    for (int j = 0; j < n; ++j) {
      double xq = xcell + dx / 2.0 * B.xq[j];
      orc_dg1d_condinit(p, xq, uinit + ((size_t)(ic - 1) * n + j) * NV);
      double w[NV] = {exp(-xq), 0, exp(-xq)};                 /* get_eq_solution :1031-1048 */
      cons(w, u_eq + ((size_t)(ic - 1) * n + j) * NV, p->gamma);
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        for (int v = 0; v < NV; ++v) {
          size_t kq = ((size_t)(ic - 1) * n + j) * NV + v, km = ((size_t)(ic - 1) * n + i) * NV + v;
          delta_u[km] = delta_u[km] + 0.5 * (uinit[kq] - u_eq[kq]) * B.P[j][i] * B.wq[j];     /* :161-164 */
        }
  }
}

/* :1749-2031 compute_update_exact_delta */
void orc_dg1d_compute_update_exact_delta(const orc_dg1d_params *p, const double *delta_u, const double *u_eq, double *dudt) {
  const int n = p->n, nx = p->nx;
  const double gamma = p->gamma;
  basis1_t B; make_basis1(n, &B);
  const double dx = p->boxlen / (double)nx, oneoverdx = 1. / dx;
  double *u_face_eq = (double *)malloc(sizeof(double) * NV * (nx + 1)), *flux_face_eq = (double *)malloc(sizeof(double) * NV * (nx + 1));
  double *flux_face = (double *)calloc(NV * (nx + 1), sizeof(double));
  double *u_left = (double *)malloc(sizeof(double) * NV * nx), *u_right = (double *)malloc(sizeof(double) * NV * nx);
  double *fv = (double *)calloc(NV * n * nx, sizeof(double)), *fve = (double *)calloc(NV * n * nx, sizeof(double));
  double *sv = (double *)calloc(NV * n * nx, sizeof(double)), *sve = (double *)calloc(NV * n * nx, sizeof(double));
#define M3(a, v, i, c) ((a)[((size_t)(c) * n + (i)) * NV + (v)])
  for (int i = 1; i <= nx + 1; ++i) {
    double xf = (double)(i - 1) * dx;
    double w[NV] = {exp(-xf), 0, exp(-xf)};
    cons(w, u_face_eq + NV * (i - 1), gamma);
    flux(u_face_eq + NV * (i - 1), flux_face_eq + NV * (i - 1), gamma);
  }
  for (int ic = 0; ic < nx; ++ic) {
    double fq[MAXN][NV], fqe[MAXN][NV], sq[MAXN][NV], sqe[MAXN][NV];
    for (int j = 0; j < n; ++j) {
      double uq[NV] = {0, 0, 0}, us[NV];
      for (int i = 0; i < n; ++i)
        for (int v = 0; v < NV; ++v) uq[v] = uq[v] + M3(delta_u, v, i, ic) * B.P[j][i];
      for (int v = 0; v < NV; ++v) us[v] = M3(u_eq, v, j, ic) + uq[v];
      flux(us, fq[j], gamma);
      flux(&M3(u_eq, 0, j, ic), fqe[j], gamma);
      source_term(us, sq[j], gamma);
      source_term(&M3(u_eq, 0, j, ic), sqe[j], gamma);
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        for (int v = 0; v < NV; ++v) {
          M3(fv, v, i, ic) = M3(fv, v, i, ic) + fq[j][v] * B.dP[j][i] * B.wq[j];
          M3(fve, v, i, ic) = M3(fve, v, i, ic) + fqe[j][v] * B.dP[j][i] * B.wq[j];
          if (p->source == 2) {
            M3(sv, v, i, ic) = M3(sv, v, i, ic) + sq[j][v] * B.P[j][i] * B.wq[j] * 0.5;
            M3(sve, v, i, ic) = M3(sve, v, i, ic) + sqe[j][v] * B.P[j][i] * B.wq[j] * 0.5;
          }
        }
    double dl[NV] = {0, 0, 0}, dr[NV] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v) {
        dl[v] = dl[v] + M3(delta_u, v, i, ic) * B.Em[i];
        dr[v] = dr[v] + M3(delta_u, v, i, ic) * B.Ep[i];
      }
    for (int v = 0; v < NV; ++v) {
      u_left[NV * ic + v] = u_face_eq[NV * ic + v] + dl[v];
      u_right[NV * ic + v] = u_face_eq[NV * (ic + 1) + v] + dr[v];
    }
  }
  /* faces 2..nx; faces 1 and nx+1 read out of bounds in the reference (bc=5) and only feed the zeroed cells */
  for (int iface = 2; iface <= nx; ++iface) {
    if (p->riemann == 1) riemann_llf(u_right + NV * (iface - 2), u_left + NV * (iface - 1), flux_face + NV * (iface - 1), gamma);
    else riemann_hllc(u_right + NV * (iface - 2), u_left + NV * (iface - 1), flux_face + NV * (iface - 1), gamma);
  }
  for (int ic = 0; ic < nx; ++ic)
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v)
        M3(dudt, v, i, ic) = oneoverdx * M3(fv, v, i, ic) - oneoverdx * M3(fve, v, i, ic)
                             - oneoverdx * (flux_face[NV * (ic + 1) + v] * B.Ep[i] - flux_face[NV * ic + v] * B.Em[i])
                             + oneoverdx * (flux_face_eq[NV * (ic + 1) + v] * B.Ep[i] - flux_face_eq[NV * ic + v] * B.Em[i])
                             + M3(sv, v, i, ic) - M3(sve, v, i, ic);
  for (int i = 0; i < n; ++i)
    for (int v = 0; v < NV; ++v) { M3(dudt, v, i, 0) = 0; M3(dudt, v, i, nx - 1) = 0; }
  free(u_face_eq); free(flux_face_eq); free(flux_face); free(u_left); free(u_right); free(fv); free(fve); free(sv); free(sve);
}

/* :1136-1152 compute_max_speed: FIRST NODE of every cell of a nodal field */
void orc_dg1d_compute_max_speed(const orc_dg1d_params *p, const double *u_nodes, double *cmax) {
  double c = 0.0;
  for (int ic = 0; ic < p->nx; ++ic) c = fmax(c, speed(u_nodes + (size_t)ic * p->n * NV, p->gamma));
  *cmax = c;
}

/* :313-330 nodal reconstruction (legendre is evaluated at the PHYSICAL coordinate, clamped to [-1,1], as shipped) */
void orc_dg1d_reconstruct(const orc_dg1d_params *p, const double *delta_u, const double *u_eq, double *uinit) {
  const int n = p->n, nx = p->nx;
  basis1_t B; make_basis1(n, &B);
  const double dx = p->boxlen / (double)nx;
  for (int ic = 1; ic <= nx; ++ic) {
    double xcell = ((double)ic - 0.5) * dx;
    for (int i = 0; i < n; ++i) {
      double xq = xcell + dx / 2.0 * B.xq[i];
      for (int v = 0; v < NV; ++v) {
        double a = 0.0;
        for (int m = 0; m < n; ++m) a = a + M3(delta_u, v, m, ic - 1) * orc_dg1d_legendre(&xq, m);
        M3(uinit, v, i, ic - 1) = M3(u_eq, v, i, ic - 1) + a;
      }
    }
  }
}

/* main loop :173-336 with integrator == 'RKi' (:282-305) */
void orc_dg1d_evolve_rki(const orc_dg1d_params *p, double *delta_u, const double *u_eq, double *uinit, double tend,
                         int max_iter, int *iters, double *t_out, double *dt_out) {
  const size_t N = (size_t)NV * p->n * p->nx;
  const double dx = p->boxlen / (double)p->nx;
  double *dudt = (double *)malloc(sizeof(double) * N), *w1 = (double *)malloc(sizeof(double) * N), *w2 = (double *)malloc(sizeof(double) * N);
  double *w3 = (double *)malloc(sizeof(double) * N), *w4 = (double *)malloc(sizeof(double) * N);
  double t = 0, dt = 0, cmax;
  int iter = 0;
  while (t < tend && (max_iter < 0 || iter < max_iter)) {
    orc_dg1d_compute_max_speed(p, uinit, &cmax);
    dt = (double)0.9f * dx / cmax / (2.0 * (double)p->n + 1.0);
    orc_dg1d_compute_update_exact_delta(p, delta_u, u_eq, dudt);
    for (size_t k = 0; k < N; ++k) w1[k] = delta_u[k] + F32(0.391752226571890) * dt * dudt[k];
    orc_dg1d_compute_update_exact_delta(p, w1, u_eq, dudt);
    for (size_t k = 0; k < N; ++k) w2[k] = F32(0.444370493651235) * delta_u[k] + F32(0.555629506348765) * w1[k] + F32(0.368410593050371) * dt * dudt[k];
    orc_dg1d_compute_update_exact_delta(p, w2, u_eq, dudt);
    for (size_t k = 0; k < N; ++k) w3[k] = F32(0.620101851488403) * delta_u[k] + F32(0.379898148511597) * w2[k] + F32(0.251891774271694) * dt * dudt[k];
    orc_dg1d_compute_update_exact_delta(p, w3, u_eq, dudt);
    for (size_t k = 0; k < N; ++k) w4[k] = F32(0.178079954393132) * delta_u[k] + F32(0.821920045606868) * w3[k] + F32(0.544974750228521) * dt * dudt[k];
    for (size_t k = 0; k < N; ++k) delta_u[k] = F32(0.517231671970585) * w2[k] + F32(0.096059710526147) * w3[k] + F32(0.063692468666290) * dt * dudt[k];
    orc_dg1d_compute_update_exact_delta(p, w4, u_eq, dudt);
    for (size_t k = 0; k < N; ++k) delta_u[k] = delta_u[k] + F32(0.386708617503269) * w4[k] + F32(0.226007483236906) * dt * dudt[k];
    orc_dg1d_reconstruct(p, delta_u, u_eq, uinit);
    t = t + dt;
    iter = iter + 1;
  }
  if (iters) *iters = iter;
  if (t_out) *t_out = t;
  if (dt_out) *dt_out = dt;
  free(dudt); free(w1); free(w2); free(w3); free(w4);
}


/* ==================================================================== plain update, limiter, 'RK1'..'RK4' */
/* :807-1028 compute_update(u,dudt) on the full state (bc 1..4; with the default bc = 5 the reference reads
 * u_right(:,0) and u_left(:,nx+1) out of bounds) */
void orc_dg1d_compute_update(const orc_dg1d_params *p, const double *u, double *dudt) {
  const int n = p->n, nx = p->nx;
  const double gamma = p->gamma;
  basis1_t B; make_basis1(n, &B);
  const double dx = p->boxlen / (double)nx, oneoverdx = 1.0 / dx;
  double *u_left = (double *)calloc(NV * nx, sizeof(double)), *u_right = (double *)calloc(NV * nx, sizeof(double));
  double *flux_face = (double *)calloc(NV * (nx + 1), sizeof(double));
  double *fv = (double *)calloc(NV * n * nx, sizeof(double)), *sv = (double *)calloc(NV * n * nx, sizeof(double));
  for (int ic = 0; ic < nx; ++ic) {
    double fq[MAXN][NV], sq[MAXN][NV];
    for (int j = 0; j < n; ++j) {
      double uq[NV] = {0, 0, 0};
      for (int i = 0; i < n; ++i)
        for (int v = 0; v < NV; ++v) uq[v] = uq[v] + M3(u, v, i, ic) * B.P[j][i];
      flux(uq, fq[j], gamma);
      source_term(uq, sq[j], gamma);
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        for (int v = 0; v < NV; ++v) {
          M3(fv, v, i, ic) = M3(fv, v, i, ic) + fq[j][v] * B.dP[j][i] * B.wq[j];
          if (p->source == 2) M3(sv, v, i, ic) = M3(sv, v, i, ic) + sq[j][v] * B.P[j][i] * B.wq[j];
        }
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v) {
        u_left[NV * ic + v] = u_left[NV * ic + v] + M3(u, v, i, ic) * B.Em[i];
        u_right[NV * ic + v] = u_right[NV * ic + v] + M3(u, v, i, ic) * B.Ep[i];
      }
  }
  for (int iface = 1; iface <= nx + 1; ++iface) {
    int ileft = iface - 1, iright = iface;
    if (p->bc == 1) { if (iface == 1) ileft = nx; if (iface == nx + 1) iright = 1; }
    if (p->bc == 2 || p->bc == 3) { if (iface == 1) ileft = 1; if (iface == nx + 1) iright = nx; }
    if (ileft < 1 || iright > nx) { if (p->bc != 4) continue; }   /* bc 5: out of bounds in the reference */
    double *ff = flux_face + NV * (iface - 1);
    void (*rs)(const double *, const double *, double *, double) = (p->riemann == 1) ? riemann_llf : riemann_hllc;
    if (ileft >= 1 && iright <= nx) rs(u_right + NV * (ileft - 1), u_left + NV * (iright - 1), ff, gamma);
    if ((p->bc == 3 || p->bc == 4) && iface == 1) {
      double t[NV];
      const double *src = (p->bc == 3) ? u_left + NV * (iright - 1) : u_right + 0;      /* bc 4: u_right(:,1) */
      t[0] = src[0]; t[1] = -src[1]; t[2] = src[2];
      rs(t, u_left + NV * (iright - 1), ff, gamma);
    }
    if ((p->bc == 3 || p->bc == 4) && iface == nx + 1) {
      double t[NV];
      const double *src = (p->bc == 3) ? u_right + NV * (ileft - 1) : u_left + NV * (nx - 1);   /* bc 4: u_left(:,nx) */
      t[0] = src[0]; t[1] = -src[1]; t[2] = src[2];
      rs(u_right + NV * (ileft - 1), t, ff, gamma);
    }
  }
  for (int ic = 0; ic < nx; ++ic)
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v)
        M3(dudt, v, i, ic) = oneoverdx * (M3(fv, v, i, ic) - (flux_face[NV * (ic + 1) + v] * B.Ep[i] - flux_face[NV * ic + v] * B.Em[i]))
                             + M3(sv, v, i, ic);
  for (int i = 0; i < n; ++i)
    for (int v = 0; v < NV; ++v) { M3(dudt, v, i, 0) = M3(dudt, v, i, 1); M3(dudt, v, i, nx - 2) = M3(dudt, v, i, nx - 1); }
  free(u_left); free(u_right); free(flux_face); free(fv); free(sv);
}

/* :1225-1233, :1250-1258, :1260-1286 */
static void cons_to_prim(const double *du, double *dw, const double *w, double gamma) {
  dw[0] = du[0];
  dw[1] = (du[1] - w[1] * du[0]) / w[0];
  dw[2] = (gamma - (double)1.0f) * (0.5 * (w[1] * w[1]) * du[0] - w[1] * du[1] + du[2]);
}
static void prim_to_cons(const double *dw, double *du, const double *w, double gamma) {
  du[0] = dw[0];
  du[1] = w[1] * dw[0] + w[0] * dw[1];
  du[2] = 0.5 * (w[1] * w[1]) * dw[0] + w[0] * w[1] * dw[1] + dw[2] / (gamma - (double)1.0f);
}
static void cons_to_char(const double *du, double *dw, const double *w, double gamma) {
  double csq = gamma * fmax(w[2], 1e-10) / fmax(w[0], 1e-10), cs = sqrt(csq), dp[NV];
  cons_to_prim(du, dp, w, gamma);
  dw[0] = dp[0] - dp[2] / csq;
  dw[1] = 0.5 * (dp[2] / csq + dp[1] * w[0] / cs);
  dw[2] = 0.5 * (dp[2] / csq - dp[1] * w[0] / cs);
}
static void char_to_cons(const double *dw, double *du, const double *w, double gamma) {
  double csq = gamma * fmax(w[2], 1e-10) / fmax(w[0], 1e-10), cs = sqrt(csq), dp[NV];
  dp[0] = dw[0] + dw[1] + dw[2];
  dp[1] = (dw[1] - dw[2]) * cs / w[0];
  dp[2] = (dw[1] + dw[2]) * csq;
  prim_to_cons(dp, du, w, gamma);
}
static double minmod3(double x, double y, double z) {      /* :1155-1165 */
  double s = copysign(1.0, x);
  if (copysign(1.0, y) == s && copysign(1.0, z) == s) return s * fmin(fmin(fabs(x), fabs(y)), fabs(z));
  return 0.0;
}
/* :414-519 limiter(u): Krivodonova moment limiter in characteristic variables (if use_limiter) + positivity fallback */
void orc_dg1d_limiter(const orc_dg1d_params *p, double *u) {
  const int n = p->n, nx = p->nx;
  const double gamma = p->gamma;
  if (n == 1) return;
  size_t N = (size_t)NV * n * nx;
  double *ul = (double *)malloc(sizeof(double) * N);
  memcpy(ul, u, sizeof(double) * N);
  if (p->use_limiter) {
    for (int ic = 1; ic <= nx; ++ic) {
      int ileft = ic - 1, iright = ic + 1;
      double switch_left = 1.0, switch_right = 1.0;
      if (p->bc == 1) { if (ic == 1) ileft = nx; if (ic == nx) iright = 1; }
      if (p->bc == 2) { if (ic == 1) ileft = 1; if (ic == nx) iright = nx; }
      if (p->bc == 3) { if (ic == 1) { ileft = 1; switch_left = -1.0; } if (ic == nx) { iright = nx; switch_right = -1.0; } }
      if (ileft < 1 || iright > nx) continue;        /* bc 4/5: out of bounds in the reference */
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
      for (int i = 1; i < n; ++i) for (int v = 0; v < NV; ++v) w_lim[i][v] = wM[i][v];
      for (int v = 0; v < NV; ++v)
        for (int i = n - 1; i >= 1; --i) {
          double w_min = minmod3(wL[i][v], wM[i][v], wR[i][v]);
          w_lim[i][v] = w_min;
          if (fabs(w_min - wM[i][v]) < (double)0.01f * fabs(wM[i][v])) break;
        }
      for (int i = n - 1; i >= 1; --i) char_to_cons(w_lim[i], &M3(ul, 0, i, ic - 1), w, gamma);
    }
  }
  for (int ic = 0; ic < nx; ++ic) {
    double w[NV], u_left[NV] = {0, 0, 0}, u_right[NV] = {0, 0, 0}, w_left[NV], w_right[NV];
    prim(&M3(ul, 0, 0, ic), w, gamma);
    for (int i = 1; i <= n; ++i)
      for (int v = 0; v < NV; ++v) {
        u_left[v] = u_left[v] + M3(ul, v, i - 1, ic) * pow((double)-1.0f, i - 1) * sqrt(2.0 * (double)i - 1.0);
        u_right[v] = u_right[v] + M3(ul, v, i - 1, ic) * sqrt(2.0 * (double)i - 1.0);
      }
    cons_to_prim(u_left, w_left, w, gamma);
    cons_to_prim(u_right, w_right, w, gamma);
    if (w_left[0] < 1e-10 || w_right[0] < 1e-10 || w_left[2] < 1e-10 || w_left[2] < 1e-10)
      for (int i = 1; i < n; ++i) for (int v = 0; v < NV; ++v) M3(ul, v, i, ic) = 0.0;
  }
  memcpy(u, ul, sizeof(double) * N);
  free(ul);
}

/* main loop :173-336 with integrator 'RK1'..'RK4' (id 1..4).  delta_u is never updated on these paths, so the nodal
 * state `uinit` that feeds the time step becomes u_eq + reconstruct(delta_u) after the first step and stays there. */
void orc_dg1d_evolve_rk(const orc_dg1d_params *p, int integrator, double *u, const double *delta_u, const double *u_eq,
                        double *uinit, double tend, int max_iter, int *iters, double *t_out, double *dt_out) {
  const size_t N = (size_t)NV * p->n * p->nx;
  const double dx = p->boxlen / (double)p->nx;
  double *dudt = (double *)malloc(sizeof(double) * N), *w1 = (double *)malloc(sizeof(double) * N), *w2 = (double *)malloc(sizeof(double) * N);
  double *w3 = (double *)malloc(sizeof(double) * N), *w4 = (double *)malloc(sizeof(double) * N);
  double t = 0, dt = 0, cmax;
  int iter = 0;
  while (t < tend && (max_iter < 0 || iter < max_iter)) {
    orc_dg1d_compute_max_speed(p, uinit, &cmax);
    dt = (double)0.9f * dx / cmax / (2.0 * (double)p->n + 1.0);
    if (integrator == 1) {
      orc_dg1d_compute_update(p, u, dudt);
      for (size_t k = 0; k < N; ++k) u[k] = u[k] + dt * dudt[k];
    } else if (integrator == 2) {
      orc_dg1d_compute_update(p, u, dudt);
      for (size_t k = 0; k < N; ++k) w1[k] = u[k] + dt * dudt[k];
      orc_dg1d_limiter(p, w1);
      orc_dg1d_compute_update(p, w1, dudt);
      for (size_t k = 0; k < N; ++k) u[k] = 0.5 * u[k] + 0.5 * w1[k] + 0.5 * dt * dudt[k];
      orc_dg1d_limiter(p, u);
    } else if (integrator == 3) {
      orc_dg1d_compute_update(p, u, dudt);
      for (size_t k = 0; k < N; ++k) w1[k] = u[k] + dt * dudt[k];
      orc_dg1d_limiter(p, w1);
      orc_dg1d_compute_update(p, w1, dudt);
      for (size_t k = 0; k < N; ++k) w2[k] = 0.75 * u[k] + 0.25 * w1[k] + 0.25 * dt * dudt[k];
      orc_dg1d_limiter(p, w2);
      orc_dg1d_compute_update(p, w2, dudt);
      for (size_t k = 0; k < N; ++k) u[k] = (double)(1.0f / 3.0f) * u[k] + (double)(2.0f / 3.0f) * w2[k] + (double)(2.0f / 3.0f) * dt * dudt[k];
      orc_dg1d_limiter(p, u);
    } else {
      for (size_t k = 0; k < N; ++k) u[k] = u[k] - u_eq[k];            /* :206 (modes minus NODAL equilibrium, as shipped) */
      orc_dg1d_compute_update(p, u, dudt);
      for (size_t k = 0; k < N; ++k) w1[k] = u[k] + F32(0.391752226571890) * dt * dudt[k];
      orc_dg1d_limiter(p, w1);
      orc_dg1d_compute_update(p, w1, dudt);
      for (size_t k = 0; k < N; ++k) w2[k] = F32(0.444370493651235) * u[k] + F32(0.555629506348765) * w1[k] + F32(0.368410593050371) * dt * dudt[k];
      orc_dg1d_limiter(p, w2);
      orc_dg1d_compute_update(p, w2, dudt);
      for (size_t k = 0; k < N; ++k) w3[k] = F32(0.620101851488403) * u[k] + F32(0.379898148511597) * w2[k] + F32(0.251891774271694) * dt * dudt[k];
      orc_dg1d_limiter(p, w3);
      orc_dg1d_compute_update(p, w3, dudt);
      for (size_t k = 0; k < N; ++k) w4[k] = F32(0.178079954393132) * u[k] + F32(0.821920045606868) * w3[k] + F32(0.544974750228521) * dt * dudt[k];
      for (size_t k = 0; k < N; ++k) u[k] = F32(0.517231671970585) * w2[k] + F32(0.096059710526147) * w3[k] + F32(0.063692468666290) * dt * dudt[k];
      orc_dg1d_limiter(p, w4);
      orc_dg1d_compute_update(p, w4, dudt);
      for (size_t k = 0; k < N; ++k) u[k] = u[k] + F32(0.386708617503269) * w4[k] + F32(0.226007483236906) * dt * dudt[k];
      orc_dg1d_limiter(p, u);
      for (size_t k = 0; k < N; ++k) u[k] = u[k] + u_eq[k];
    }
    orc_dg1d_reconstruct(p, delta_u, u_eq, uinit);
    t = t + dt;
    iter = iter + 1;
  }
  if (iters) *iters = iter;
  if (t_out) *t_out = t;
  if (dt_out) *dt_out = dt;
  free(dudt); free(w1); free(w2); free(w3); free(w4);
}

/* program dg :33-47: projection of the full initial state onto the modes `u` */
void orc_dg1d_project(const orc_dg1d_params *p, const double *u_nodes, double *u_modes) {
  const int n = p->n, nx = p->nx;
  basis1_t B; make_basis1(n, &B);
  memset(u_modes, 0, sizeof(double) * NV * n * nx);
  for (int ic = 0; ic < nx; ++ic)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        for (int v = 0; v < NV; ++v) M3(u_modes, v, i, ic) = M3(u_modes, v, i, ic) + 0.5 * M3(u_nodes, v, j, ic) * B.P[j][i] * B.wq[j];
}


/* ==================================================================== 'RKw' / 'RKe': compute_update_exact, limiter_TDV,
 * limiter_cons.  get_eq_solution :1031-1048 is w = (exp(-x), 0, exp(-x)). */
static void eq_cons(double x, double *u, double gamma) {
  double w[NV] = {exp(-x), 0, exp(-x)};
  cons(w, u, gamma);
}
/* :1380-1744 compute_update_exact(u,u_eq,dudt): full-state modes `u`, equilibrium MODES `u_eq`.  Faces 1 and nx+1 read
 * u_right(:,0) / u_left(:,nx+1) out of bounds; with bc = 4 or 5 that Riemann result is discarded and replaced by the
 * physical flux of a boundary state (:1516-1620), with any other bc it would be used -> only bc 4, 5 are defined. */
void orc_dg1d_compute_update_exact(const orc_dg1d_params *p, const double *u, const double *u_eq, double *dudt) {
  const int n = p->n, nx = p->nx;
  const double gamma = p->gamma;
  basis1_t B; make_basis1(n, &B);
  const double dx = p->boxlen / (double)nx, oneoverdx = 1. / dx;
  double *u_face_eq = (double *)malloc(sizeof(double) * NV * (nx + 1)), *flux_face_eq = (double *)malloc(sizeof(double) * NV * (nx + 1));
  double *flux_face = (double *)calloc(NV * (nx + 1), sizeof(double));
  double *u_left = (double *)malloc(sizeof(double) * NV * nx), *u_right = (double *)malloc(sizeof(double) * NV * nx);
  double *fv = (double *)calloc(NV * n * nx, sizeof(double)), *fve = (double *)calloc(NV * n * nx, sizeof(double));
  double *sv = (double *)calloc(NV * n * nx, sizeof(double)), *sve = (double *)calloc(NV * n * nx, sizeof(double));
  for (int i = 1; i <= nx + 1; ++i) {
    eq_cons((double)(i - 1) * dx, u_face_eq + NV * (i - 1), gamma);
    flux(u_face_eq + NV * (i - 1), flux_face_eq + NV * (i - 1), gamma);
  }
  for (int ic = 0; ic < nx; ++ic) {
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
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        for (int v = 0; v < NV; ++v) {
          M3(fv, v, i, ic) = M3(fv, v, i, ic) + fq[j][v] * B.dP[j][i] * B.wq[j];
          M3(fve, v, i, ic) = M3(fve, v, i, ic) + fqe[j][v] * B.dP[j][i] * B.wq[j];
          if (p->source == 2) {
            M3(sv, v, i, ic) = M3(sv, v, i, ic) + sq[j][v] * B.P[j][i] * B.wq[j];
            M3(sve, v, i, ic) = M3(sve, v, i, ic) + sqe[j][v] * B.P[j][i] * B.wq[j];
          }
        }
    double dl[NV] = {0, 0, 0}, dr[NV] = {0, 0, 0}, uul[NV], uur[NV];
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v) {
        dl[v] = dl[v] + M3(u, v, i, ic) * B.Em[i];
        dr[v] = dr[v] + M3(u, v, i, ic) * B.Ep[i];
      }
    eq_cons((double)(ic + 1) * dx, uur, gamma);      /* x_right = icell*dx */
    eq_cons((double)ic * dx, uul, gamma);            /* x_left = (icell-1)*dx */
    for (int v = 0; v < NV; ++v) {
      u_left[NV * ic + v] = u_face_eq[NV * ic + v] + (dl[v] - uul[v]);
      u_right[NV * ic + v] = u_face_eq[NV * (ic + 1) + v] + (dr[v] - uur[v]);
    }
  }
  void (*rs)(const double *, const double *, double *, double) = (p->riemann == 1) ? riemann_llf : riemann_hllc;
  for (int iface = 2; iface <= nx; ++iface) rs(u_right + NV * (iface - 2), u_left + NV * (iface - 1), flux_face + NV * (iface - 1), gamma);
  {
    double a[NV], b[NV], t[NV];
    if (p->bc == 4) {
      eq_cons((double)-0.5f * dx + dx / 2.0 * (double)(1), a, gamma);
      eq_cons((double)0.5f * dx + dx / 2.0 * (double)(-1), b, gamma);
      for (int v = 0; v < NV; ++v) t[v] = a[v] + u_left[v] - b[v];
      flux(t, flux_face, gamma);
      eq_cons((double)((float)nx + 0.5f) * dx + dx / 2.0 * (double)(-1), a, gamma);
      eq_cons((double)nx * dx, b, gamma);
      for (int v = 0; v < NV; ++v) t[v] = a[v] + u_right[NV * (nx - 1) + v] - b[v];
      flux(t, flux_face + NV * nx, gamma);
    } else {                                               /* bc 5 "mimic FVM": the MEAN MODE of the end cell (sic) */
      eq_cons((double)-0.5f * dx, a, gamma);
      eq_cons((double)0.5f * dx, b, gamma);
      for (int v = 0; v < NV; ++v) t[v] = a[v] + M3(u, v, 0, 0) - b[v];
      flux(t, flux_face, gamma);
      eq_cons((double)((float)nx + 0.5f) * dx, a, gamma);
      eq_cons((double)((float)nx - 0.5f) * dx, b, gamma);
      for (int v = 0; v < NV; ++v) t[v] = a[v] + M3(u, v, 0, nx - 1) - b[v];
      flux(t, flux_face + NV * nx, gamma);
    }
  }
  for (int ic = 0; ic < nx; ++ic)
    for (int i = 0; i < n; ++i)
      for (int v = 0; v < NV; ++v)
        M3(dudt, v, i, ic) = oneoverdx * M3(fv, v, i, ic) - oneoverdx * M3(fve, v, i, ic)
                             - oneoverdx * (flux_face[NV * (ic + 1) + v] * B.Ep[i] - flux_face[NV * ic + v] * B.Em[i])
                             + oneoverdx * (flux_face_eq[NV * (ic + 1) + v] * B.Ep[i] - flux_face_eq[NV * ic + v] * B.Em[i])
                             + M3(sv, v, i, ic) - M3(sve, v, i, ic);
  free(u_face_eq); free(flux_face_eq); free(flux_face); free(u_left); free(u_right); free(fv); free(fve); free(sv); free(sve);
}

/* :520-600 limiter_TDV(u).  Its moment-limiting block indexes the neighbours with the loop variable of ANOTHER loop
 * (`u(1:nvar,1,i-1)`, :550-552: undefined behaviour), so only use_limiter = .false. (the shipped value) is defined:
 * what remains is the positivity fallback on the TRACES of the conserved variables (no cons_to_prim here). */
void orc_dg1d_limiter_tdv(const orc_dg1d_params *p, double *u) {
  const int n = p->n, nx = p->nx;
  if (n == 1) return;
  for (int ic = 0; ic < nx; ++ic) {
    double ul[NV] = {0, 0, 0}, ur[NV] = {0, 0, 0};
    for (int i = 1; i <= n; ++i)
      for (int v = 0; v < NV; ++v) {
        ul[v] = ul[v] + M3(u, v, i - 1, ic) * pow((double)-1.0f, i - 1) * sqrt(2.0 * (double)i - 1.0);
        ur[v] = ur[v] + M3(u, v, i - 1, ic) * sqrt(2.0 * (double)i - 1.0);
      }
    if (ul[0] < 1e-10 || ur[0] < 1e-10 || ul[2] < 1e-10 || ul[2] < 1e-10)
      for (int i = 1; i < n; ++i) for (int v = 0; v < NV; ++v) M3(u, v, i, ic) = 0.0;
  }
}

/* :602-734 limiter_cons(u): the moment limiter of limiter() applied to the conserved moments directly */
void orc_dg1d_limiter_cons(const orc_dg1d_params *p, double *u) {
  const int n = p->n, nx = p->nx;
  const double gamma = p->gamma;
  if (n == 1) return;
  size_t N = (size_t)NV * n * nx;
  double *ul = (double *)malloc(sizeof(double) * N);
  memcpy(ul, u, sizeof(double) * N);
  if (p->use_limiter) {
    for (int ic = 1; ic <= nx; ++ic) {
      int ileft = ic - 1, iright = ic + 1;
      double switch_left = 1.0, switch_right = 1.0;
      if (p->bc == 1) { if (ic == 1) ileft = nx; if (ic == nx) iright = 1; }
      if (p->bc == 2 || p->bc == 4) { if (ic == 1) ileft = 1; if (ic == nx) iright = nx; }
      if (p->bc == 3) { if (ic == 1) { ileft = 1; switch_left = -1.0; } if (ic == nx) { iright = nx; switch_right = -1.0; } }
      if (ileft < 1 || iright > nx) continue;        /* bc 5: out of bounds in the reference */
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
      for (int i = 1; i < n; ++i) for (int v = 0; v < NV; ++v) w_lim[i][v] = wM[i][v];
      for (int v = 0; v < NV; ++v)
        for (int i = n - 1; i >= 1; --i) {
          double w_min = minmod3(wL[i][v], wM[i][v], wR[i][v]);
          w_lim[i][v] = w_min;
          if (fabs(w_min - wM[i][v]) < (double)0.01f * fabs(wM[i][v])) break;
        }
      for (int i = n - 1; i >= 1; --i) for (int v = 0; v < NV; ++v) M3(ul, v, i, ic - 1) = w_lim[i][v];
    }
  }
  for (int ic = 0; ic < nx; ++ic) {
    double w[NV], u_left[NV] = {0, 0, 0}, u_right[NV] = {0, 0, 0}, w_left[NV], w_right[NV];
    prim(&M3(ul, 0, 0, ic), w, gamma);
    for (int i = 1; i <= n; ++i)
      for (int v = 0; v < NV; ++v) {
        u_left[v] = u_left[v] + M3(ul, v, i - 1, ic) * pow((double)-1.0f, i - 1) * sqrt(2.0 * (double)i - 1.0);
        u_right[v] = u_right[v] + M3(ul, v, i - 1, ic) * sqrt(2.0 * (double)i - 1.0);
      }
    cons_to_prim(u_left, w_left, w, gamma);
    cons_to_prim(u_right, w_right, w, gamma);
    if (w_left[0] < 1e-10 || w_right[0] < 1e-10 || w_left[2] < 1e-10 || w_left[2] < 1e-10)
      for (int i = 1; i < n; ++i) for (int v = 0; v < NV; ++v) M3(ul, v, i, ic) = 0.0;
  }
  memcpy(u, ul, sizeof(double) * N);
  free(ul);
}

/* main loop :173-336 with integrator 'RKw' (5, :229-270: SSPRK(5,4) on the full state, limiter_TDV on u - u_eq_modes
 * after every stage; delta_u ends as u - u_eq_modes and feeds `uinit`) or 'RKe' (6, :273-280: RK2 on the perturbation
 * with limiter_cons; u_eq are the NODAL equilibrium values as in 'RKi'). */
void orc_dg1d_evolve_w(const orc_dg1d_params *p, int integrator, double *u, double *delta_u, const double *u_eq_nodes,
                       const double *u_eq_modes, double *uinit, double tend, int max_iter, int *iters, double *t_out, double *dt_out) {
  const size_t N = (size_t)NV * p->n * p->nx;
  const double dx = p->boxlen / (double)p->nx;
  double *dudt = (double *)malloc(sizeof(double) * N), *w1 = (double *)malloc(sizeof(double) * N), *w2 = (double *)malloc(sizeof(double) * N);
  double *w3 = (double *)malloc(sizeof(double) * N), *w4 = (double *)malloc(sizeof(double) * N);
  double t = 0, dt = 0, cmax;
  int iter = 0;
  const double *q = u_eq_modes;
#define LIM_W(w) do { for (size_t k = 0; k < N; ++k) delta_u[k] = (w)[k] - q[k]; orc_dg1d_limiter_tdv(p, delta_u); \
                      for (size_t k = 0; k < N; ++k) (w)[k] = q[k] + delta_u[k]; } while (0)
  while (t < tend && (max_iter < 0 || iter < max_iter)) {
    orc_dg1d_compute_max_speed(p, uinit, &cmax);
    dt = (double)0.9f * dx / cmax / (2.0 * (double)p->n + 1.0);
    if (integrator == 5) {
      orc_dg1d_compute_update_exact(p, u, q, dudt);
      for (size_t k = 0; k < N; ++k) w1[k] = u[k] + F32(0.391752226571890) * dt * dudt[k];
      LIM_W(w1);
      orc_dg1d_compute_update_exact(p, w1, q, dudt);
      for (size_t k = 0; k < N; ++k) w2[k] = F32(0.444370493651235) * u[k] + F32(0.555629506348765) * w1[k] + F32(0.368410593050371) * dt * dudt[k];
      LIM_W(w2);
      orc_dg1d_compute_update_exact(p, w2, q, dudt);
      for (size_t k = 0; k < N; ++k) w3[k] = F32(0.620101851488403) * u[k] + F32(0.379898148511597) * w2[k] + F32(0.251891774271694) * dt * dudt[k];
      LIM_W(w3);
      orc_dg1d_compute_update_exact(p, w3, q, dudt);
      for (size_t k = 0; k < N; ++k) w4[k] = F32(0.178079954393132) * u[k] + F32(0.821920045606868) * w3[k] + F32(0.544974750228521) * dt * dudt[k];
      for (size_t k = 0; k < N; ++k) u[k] = F32(0.517231671970585) * w2[k] + F32(0.096059710526147) * w3[k] + F32(0.063692468666290) * dt * dudt[k];
      LIM_W(w4);                                                     /* :262 writes delta_u + u_eq_modes: the sum commutes */
      orc_dg1d_compute_update_exact(p, w4, q, dudt);
      for (size_t k = 0; k < N; ++k) u[k] = u[k] + F32(0.386708617503269) * w4[k] + F32(0.226007483236906) * dt * dudt[k];
      LIM_W(u);
    } else {
      orc_dg1d_compute_update_exact_delta(p, delta_u, u_eq_nodes, dudt);
      orc_dg1d_limiter_cons(p, delta_u);
      for (size_t k = 0; k < N; ++k) w1[k] = delta_u[k] + dt * dudt[k];
      orc_dg1d_compute_update_exact_delta(p, w1, u_eq_nodes, dudt);
      orc_dg1d_limiter_cons(p, w1);
      for (size_t k = 0; k < N; ++k) delta_u[k] = 0.5 * delta_u[k] + 0.5 * w1[k] + 0.5 * dt * dudt[k];
    }
    orc_dg1d_reconstruct(p, delta_u, u_eq_nodes, uinit);
    t = t + dt;
    iter = iter + 1;
  }
#undef LIM_W
  if (iters) *iters = iter;
  if (t_out) *t_out = t;
  if (dt_out) *dt_out = dt;
  free(dudt); free(w1); free(w2); free(w3); free(w4);
}
