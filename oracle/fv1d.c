/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement (plain C, FP64, -ffp-contract=off) of the 1D finite-volume paths of the reference:
 *   fvm.f90          (module fvm_commons.f90)  plain first-order FV + centred gravity source, SSP-RK2
 *   benchmark_1d.f90 (module parameters.f90)   three schemes: 'FVM' plain, 'EQL' equilibrium subtraction,
 *                                              'WB1' local hydrostatic reconstruction (default), SSP-RK2
 * PARITY PINNED TO THE REFERENCE'S OWN SOURCE TEXT: fvm.f90 and benchmark_1d.f90 are EXECUTED, unmodified, by the
 * Fortran-90 interpreter oracle/f90interp.py (no Fortran compiler in the image); this file reproduces the vectors
 * (tests/golden/ref_fv1d.npz, generator tests/golden/make_ref_golden.py) BIT FOR BIT: initial conditions, the update
 * routines of all schemes, max speed and whole time loops (tests/test_reference_pins.py).  Additional pins: the
 * invariants in tests/test_oracle_fv1d.py (EQL: bitwise-zero RHS at the discrete equilibrium; WB1: isentropic
 * equilibrium preserved to round-off, isothermal to O(dx^2); first-order convergence).
 *
 * Layout: Fortran u(nvar,nx) == C double[nx][3].  Literal kinds as in the reference (real(4) literals promoted).
 * Debug `write`s are omitted.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NV 3

/* ============================================================================================ fvm.f90 */
typedef struct {
  int nx;        /* fvm_commons.f90:6 */
  int bc;        /* :13  1 periodic, 2 zero gradient */
  int source;    /* :14  1 none, 2 gravity */
  int n;         /* :4   only used in dt = 0.8*dx/cmax/(2n+1) */
  double gamma;  /* :17 */
  double boxlen; /* :16 */
} orc_fvm1d_params;

/* fvm.f90:176-186 */
static void fvm_prim(const double *u, double *w, double gamma) {
  w[0] = u[0];
  w[1] = u[1] / w[0];
  w[2] = (gamma - (double)1.0f) * (u[2] - 0.5 * w[0] * (w[1] * w[1]));
}
/* fvm.f90:300-312 */
static double fvm_speed(const double *u, double gamma) {
  double w[NV];
  fvm_prim(u, w, gamma);
  double cs = sqrt(gamma * fmax(w[2], 1e-10) / fmax(w[0], 1e-10));
  return fabs(w[1]) + cs;
}
/* fvm.f90:284-298 */
static void fvm_flux(const double *u, double *f, double gamma) {
  double w[NV];
  fvm_prim(u, w, gamma);
  f[0] = w[1] * u[0];
  f[1] = w[1] * u[1] + w[2];
  f[2] = w[1] * u[2] + w[2] * w[1];
}
/* fvm.f90:270-283 */
static void fvm_llflux(const double *ul, const double *ur, double *fg, double gamma) {
  double cl = fvm_speed(ul, gamma), cr = fvm_speed(ur, gamma);
  double cmax = fmax(cl, cr), fl[NV], fr[NV];
  fvm_flux(ul, fl, gamma);
  fvm_flux(ur, fr, gamma);
  for (int v = 0; v < NV; ++v) fg[v] = 0.5 * (fr[v] + fl[v]) - 0.5 * cmax * (ur[v] - ul[v]);
}
/* fvm.f90:104-174 condinit (conservative values at x) */
void orc_fvm1d_condinit(const orc_fvm1d_params *p, int ninit, double x, double *uu) {
  const double dpi = acos(-1.0), gamma = p->gamma;
  double ww[NV] = {0, 0, 0};
  switch (ninit) {
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
    default: /* 7 */
      ww[0] = (1 - (gamma - 1) / (gamma) * x); ww[1] = 0; ww[2] = pow(ww[0], gamma);
      break;
  }
  uu[0] = ww[0];
  uu[1] = ww[0] * ww[1];
  uu[2] = ww[2] / (gamma - (double)1.0f) + 0.5 * ww[0] * (ww[1] * ww[1]);
}
void orc_fvm1d_initial_conditions(const orc_fvm1d_params *p, int ninit, double *u) {
  double dx = p->boxlen / (double)p->nx;
  for (int i = 1; i <= p->nx; ++i) orc_fvm1d_condinit(p, ninit, ((double)i - 0.5) * dx, u + NV * (i - 1));
}
/* fvm.f90:314-330 */
void orc_fvm1d_compute_max_speed(const orc_fvm1d_params *p, const double *u, double *cmax) {
  double c = 0.0;
  for (int i = 0; i < p->nx; ++i) c = fmax(c, fvm_speed(u + NV * i, p->gamma));
  *cmax = c;
}
/* fvm.f90:188-251 compute_update */
void orc_fvm1d_compute_update(const orc_fvm1d_params *p, const double *u, double *dudt) {
  const int nx = p->nx;
  const double dx = p->boxlen / (double)nx, oneoverdx = 1.0 / dx, gamma = p->gamma;
  double *fl = (double *)malloc(sizeof(double) * NV * nx), *fr = (double *)malloc(sizeof(double) * NV * nx);
  double *src = (double *)calloc(NV * nx, sizeof(double));
  for (int iface = 1; iface <= nx; ++iface) {
    int um = iface - 1, uc = iface, up = iface + 1;
    if (p->bc == 1) { if (iface == 1) um = nx; if (iface == nx) up = 1; }
    if (p->bc == 2) { if (iface == 1) um = 1; if (iface == nx) up = nx; }
    fvm_llflux(u + NV * (um - 1), u + NV * (uc - 1), fl + NV * (iface - 1), gamma);
    fvm_llflux(u + NV * (uc - 1), u + NV * (up - 1), fr + NV * (iface - 1), gamma);
  }
  if (p->source == 2) {
    for (int ic = 1; ic <= nx; ++ic) {
      double x_minus = ((double)(ic - 1) - 0.5) * dx, x_plus = ((double)(ic + 1) - 0.5) * dx;
      if (ic == 1) x_minus = ((double)1 - 0.5) * dx;
      if (ic == nx) x_plus = ((double)nx - 0.5) * dx;
      double w[NV];
      fvm_prim(u + NV * (ic - 1), w, gamma);           /* compute_source :253-264 */
      src[NV * (ic - 1) + 0] = 0;
      src[NV * (ic - 1) + 1] = -w[0] * 1 * (x_plus - x_minus) / (2 * dx);
      src[NV * (ic - 1) + 2] = -w[0] * w[1] * 1 * (x_plus - x_minus) / (2 * dx);
    }
  }
  for (int i = 0; i < nx; ++i)
    for (int v = 0; v < NV; ++v) dudt[NV * i + v] = -oneoverdx * (fr[NV * i + v] - fl[NV * i + v]) + src[NV * i + v];
  free(fl); free(fr); free(src);
}
/* fvm.f90:56-76 main loop */
void orc_fvm1d_evolve(const orc_fvm1d_params *p, double *u, double tend, int max_iter, int *iters, double *t_out,
                      double *dt_out) {
  const int n = NV * p->nx;
  const double dx = p->boxlen / (double)p->nx;
  double *dudt = (double *)malloc(sizeof(double) * n), *w1 = (double *)malloc(sizeof(double) * n);
  double t = 0, dt = 0, cmax;
  int iter = 0;
  while (t < tend && (max_iter < 0 || iter < max_iter)) {
    orc_fvm1d_compute_max_speed(p, u, &cmax);
    dt = (double)0.8f * dx / cmax / (2.0 * (double)p->n + 1.0);
    orc_fvm1d_compute_update(p, u, dudt);
    for (int k = 0; k < n; ++k) w1[k] = u[k] + dt * dudt[k];
    orc_fvm1d_compute_update(p, w1, dudt);
    for (int k = 0; k < n; ++k) u[k] = 0.5 * u[k] + 0.5 * w1[k] + 0.5 * dt * dudt[k];
    t = t + dt;
    iter = iter + 1;
  }
  if (iters) *iters = iter;
  if (t_out) *t_out = t;
  if (dt_out) *dt_out = dt;
  free(dudt); free(w1);
}

/* ============================================================================================ benchmark_1d.f90 */
typedef struct {
  int nx;            /* parameters.f90:3 */
  int bc;            /* :12  1 periodic, 2 zero gradient, 3 reflexive */
  int nequilibrium;  /* :13 */
  int solver;        /* :8   1 'FVM', 2 'EQL', 3 'WB1' */
  double gamma;      /* :17 */
  double boxlen;     /* :16 */
} orc_fv1d_params;

/* benchmark_1d.f90:103-112 / :114-123 */
static void b1_prim(const orc_fv1d_params *p, const double *u, double *w) {
  w[0] = u[0];
  w[1] = u[1] / w[0];
  w[2] = (p->gamma - (double)1.0f) * (u[2] - 0.5 * w[0] * (w[1] * w[1]));
}
static void b1_cons(const orc_fv1d_params *p, const double *w, double *u) {
  u[0] = w[0];
  u[1] = w[0] * w[1];
  u[2] = w[2] / (p->gamma - (double)1.0f) + 0.5 * w[0] * (w[1] * w[1]);
}
/* :125-152 get_equilibrium_solution (primitives) */
static void b1_eq(const orc_fv1d_params *p, double x, double *w) {
  const double gamma = p->gamma;
  if (p->nequilibrium == 3) {
    double base = (1 - ((gamma - 1) / gamma) * 1 * x);
    w[0] = pow(base, (1 / (gamma - 1)));
    w[1] = 0;
    w[2] = pow(base, (gamma / (gamma - 1)));
  } else {
    w[0] = exp(-x); w[1] = 0; w[2] = exp(-x);
  }
}
void orc_fv1d_get_x(const orc_fv1d_params *p, double *x) {
  double dx = p->boxlen / (double)p->nx;
  for (int i = 1; i <= p->nx; ++i) x[i - 1] = (double)((float)i - 0.5f) * dx;
}
void orc_fv1d_get_equilibrium_solution(const orc_fv1d_params *p, const double *x, double *w, int size) {
  for (int i = 0; i < size; ++i) b1_eq(p, x[i], w + NV * i);
}
/* :37-63 get_initial_conditions */
void orc_fv1d_get_initial_conditions(const orc_fv1d_params *p, int ninit, double eta, const double *x, double *u) {
  const double gamma = p->gamma;
  for (int i = 0; i < p->nx; ++i) {
    double w[NV], xx = x[i];
    if (ninit == 3) {
      double base = (1 - ((gamma - 1) / gamma) * 1 * xx);
      w[0] = pow(base, (1 / (gamma - 1))); w[1] = 0; w[2] = pow(base, (gamma / (gamma - 1)));
    } else {
      w[0] = exp(-xx); w[1] = 0; w[2] = exp(-xx);
      if (ninit == 2) { double d = xx - p->boxlen / 2.; w[2] = w[2] + eta * exp(-100 * (d * d)); }
    }
    b1_cons(p, w, u + NV * i);
  }
}
/* :167-179 */
static double b1_speed(const orc_fv1d_params *p, const double *u) {
  double w[NV];
  b1_prim(p, u, w);
  double cs = sqrt(p->gamma * fmax(w[2], 1e-10) / fmax(w[0], 1e-10));
  return fabs(w[1]) + cs;
}
void orc_fv1d_compute_max_speed(const orc_fv1d_params *p, const double *u, double *cmax) {
  double c = 0.0;
  for (int i = 0; i < p->nx; ++i) c = fmax(c, b1_speed(p, u + NV * i));
  *cmax = c;
}
/* :181-198 */
static void b1_flux(const orc_fv1d_params *p, const double *u, double *f) {
  double w[NV];
  b1_prim(p, u, w);
  f[0] = w[1] * u[0];
  f[1] = w[1] * u[1] + w[2];
  f[2] = w[1] * u[2] + w[2] * w[1];
}
/* :437-451 */
static void b1_llflux(const orc_fv1d_params *p, const double *ul, const double *ur, const double *fl, const double *fr,
                      double *fg) {
  double cmax = fmax(b1_speed(p, ul), b1_speed(p, ur));
  for (int v = 0; v < NV; ++v) fg[v] = 0.5 * (fr[v] + fl[v]) - 0.5 * cmax * (ur[v] - ul[v]);
}
/* :380-406 get_source: note that the callers pass CONSERVATIVE variables as `w` (:365-366, :541) */
static void b1_get_source(const double *w, double *s, const double *x, int size) {
  double delta = 1 / (double)size;
  for (int i = 0; i < size; ++i) {
    double xm = (i == 0) ? x[0] - delta : x[i - 1];
    double xp = (i == size - 1) ? x[size - 1] + delta : x[i + 1];
    s[NV * i + 0] = 0;
    s[NV * i + 1] = -w[NV * i + 0] * 1 * (xp - xm) / (2 * delta);
    s[NV * i + 2] = -w[NV * i + 0] * w[NV * i + 1] * 1 * (xp - xm) / (2 * delta);
  }
}
static void face_indices(const orc_fv1d_params *p, int iface, int *il, int *ir) {
  const int nx = p->nx;
  int ileft = iface - 1, iright = iface;
  if (p->bc == 1) { if (iface == 1) ileft = nx; if (iface == nx + 1) iright = 1; }
  if (p->bc == 2 || p->bc == 3) { if (iface == 1) ileft = 1; if (iface == nx + 1) iright = nx; }
  *il = ileft; *ir = iright;
}
/* the reflexive-boundary face overrides shared by compute_update (:330-361) and compute_update_fvm (:516-538) */
static void bc3_faces(const orc_fv1d_params *p, const double *delta_w, const double *u_left, const double *u_right,
                      const double *f_left, const double *f_right, double *flux_riemann) {
  const int nx = p->nx;
  if (p->bc != 3) return;
  {
    double w_minus[NV] = {1., 0., 1.}, u_face[NV], a[NV], f_minus[NV];
    b1_cons(p, w_minus, u_face);
    for (int v = 0; v < NV; ++v) a[v] = u_face[v] + delta_w[v];
    b1_flux(p, a, f_minus);
    b1_llflux(p, a, u_left + NV * 0, f_minus, f_left + NV * 0, flux_riemann + NV * 0);   /* iface = 1, iright = 1 */
  }
  {
    double w_plus[NV] = {1., 0., 1.}, u_plus[NV], a[NV], b[NV], f_plus[NV];
    b1_cons(p, w_plus, u_plus);
    for (int v = 0; v < NV; ++v) a[v] = u_plus[v] + delta_w[NV * (nx - 1) + v];
    b1_flux(p, a, f_plus);
    for (int v = 0; v < NV; ++v) b[v] = u_right[NV * (nx - 1) + v] + delta_w[NV * (nx - 1) + v];   /* ileft = nx */
    b1_llflux(p, b, u_plus, f_right + NV * (nx - 1), f_plus, flux_riemann + NV * nx);
  }
}

/* :263-377 compute_update ('EQL'): equilibrium subtraction */
void orc_fv1d_compute_update(const orc_fv1d_params *p, const double *u, const double *w_eq, double *dudt) {
  const int nx = p->nx, nf = nx + 1;
  const double dx = p->boxlen / (double)nx, oneoverdx = 1 / dx;
  double *u_eq = (double *)malloc(sizeof(double) * NV * nx), *delta_w = (double *)malloc(sizeof(double) * NV * nx);
  double *x = (double *)malloc(sizeof(double) * nx), *xf = (double *)malloc(sizeof(double) * nf);
  double *w_eq_f = (double *)malloc(sizeof(double) * NV * nf), *u_eq_f = (double *)malloc(sizeof(double) * NV * nf);
  double *u_left = (double *)malloc(sizeof(double) * NV * nx), *u_right = (double *)malloc(sizeof(double) * NV * nx);
  double *f_left = (double *)malloc(sizeof(double) * NV * nx), *f_right = (double *)malloc(sizeof(double) * NV * nx);
  double *flux_eq = (double *)malloc(sizeof(double) * NV * nf), *fr = (double *)malloc(sizeof(double) * NV * nf);
  double *s = (double *)malloc(sizeof(double) * NV * nx), *s_eq = (double *)malloc(sizeof(double) * NV * nx);
  for (int i = 0; i < nx; ++i) b1_cons(p, w_eq + NV * i, u_eq + NV * i);
  for (int k = 0; k < NV * nx; ++k) delta_w[k] = u[k] - u_eq[k];
  for (int i = 1; i <= nx; ++i) x[i - 1] = (double)((float)i - 0.5f) * dx;
  for (int i = 1; i <= nf; ++i) xf[i - 1] = (double)(i - 1) * dx;
  for (int i = 0; i < nf; ++i) { b1_eq(p, xf[i], w_eq_f + NV * i); b1_cons(p, w_eq_f + NV * i, u_eq_f + NV * i); }
  for (int i = 0; i < nx; ++i)
    for (int v = 0; v < NV; ++v) {
      u_left[NV * i + v] = delta_w[NV * i + v] + u_eq_f[NV * i + v];
      u_right[NV * i + v] = delta_w[NV * i + v] + u_eq_f[NV * (i + 1) + v];
    }
  for (int i = 0; i < nx; ++i) { b1_flux(p, u_left + NV * i, f_left + NV * i); b1_flux(p, u_right + NV * i, f_right + NV * i); }
  for (int i = 0; i < nf; ++i) b1_flux(p, u_eq_f + NV * i, flux_eq + NV * i);
  for (int iface = 1; iface <= nf; ++iface) {
    int il, ir;
    face_indices(p, iface, &il, &ir);
    b1_llflux(p, u_right + NV * (il - 1), u_left + NV * (ir - 1), f_right + NV * (il - 1), f_left + NV * (ir - 1),
              fr + NV * (iface - 1));
  }
  bc3_faces(p, delta_w, u_left, u_right, f_left, f_right, fr);
  b1_get_source(u, s, x, nx);
  b1_get_source(u_eq, s_eq, x, nx);
  for (int i = 0; i < nx; ++i)
    for (int v = 0; v < NV; ++v)
      dudt[NV * i + v] = -(fr[NV * (i + 1) + v] - fr[NV * i + v]) * oneoverdx + s[NV * i + v] +
                         (flux_eq[NV * (i + 1) + v] - flux_eq[NV * i + v]) * oneoverdx - s_eq[NV * i + v];
  for (int v = 0; v < NV; ++v) { dudt[v] = dudt[NV + v]; dudt[NV * (nx - 1) + v] = dudt[NV * (nx - 2) + v]; }
  free(u_eq); free(delta_w); free(x); free(xf); free(w_eq_f); free(u_eq_f); free(u_left); free(u_right);
  free(f_left); free(f_right); free(flux_eq); free(fr); free(s); free(s_eq);
}

/* :454-549 compute_update_fvm ('FVM'): plain scheme */
void orc_fv1d_compute_update_fvm(const orc_fv1d_params *p, const double *u, const double *w_eq, double *dudt) {
  const int nx = p->nx, nf = nx + 1;
  const double dx = p->boxlen / (double)nx, oneoverdx = 1 / dx;
  double *u_eq = (double *)malloc(sizeof(double) * NV * nx), *delta_w = (double *)malloc(sizeof(double) * NV * nx);
  double *x = (double *)malloc(sizeof(double) * nx);
  double *f = (double *)malloc(sizeof(double) * NV * nx), *fr = (double *)malloc(sizeof(double) * NV * nf);
  double *s = (double *)malloc(sizeof(double) * NV * nx);
  for (int i = 0; i < nx; ++i) b1_cons(p, w_eq + NV * i, u_eq + NV * i);
  for (int k = 0; k < NV * nx; ++k) delta_w[k] = u[k] - u_eq[k];
  for (int i = 1; i <= nx; ++i) x[i - 1] = (double)((float)i - 0.5f) * dx;
  for (int i = 0; i < nx; ++i) b1_flux(p, u + NV * i, f + NV * i);
  for (int iface = 1; iface <= nf; ++iface) {
    int il, ir;
    face_indices(p, iface, &il, &ir);
    b1_llflux(p, u + NV * (il - 1), u + NV * (ir - 1), f + NV * (il - 1), f + NV * (ir - 1), fr + NV * (iface - 1));
  }
  bc3_faces(p, delta_w, u, u, f, f, fr);
  b1_get_source(u, s, x, nx);
  for (int i = 0; i < nx; ++i)
    for (int v = 0; v < NV; ++v) dudt[NV * i + v] = -(fr[NV * (i + 1) + v] - fr[NV * i + v]) * oneoverdx + s[NV * i + v];
  for (int v = 0; v < NV; ++v) { dudt[v] = dudt[NV + v]; dudt[NV * (nx - 1) + v] = dudt[NV * (nx - 2) + v]; }
  free(u_eq); free(delta_w); free(x); free(f); free(fr); free(s);
}

/* :553-747 compute_update_sr ('WB1'): local hydrostatic reconstruction; phi(x) = x (:749-755) */
void orc_fv1d_compute_update_sr(const orc_fv1d_params *p, const double *u, const double *w_eq, double *dudt) {
  (void)w_eq;
  const int nx = p->nx, nf = nx + 1;
  const double gamma = p->gamma;
  const double dx = p->boxlen / (double)nx, oneoverdx = p->boxlen / dx;   /* sic :574 */
  double *x = (double *)malloc(sizeof(double) * nx), *xf = (double *)malloc(sizeof(double) * nf);
  double *w = (double *)malloc(sizeof(double) * NV * nx);
  double *w_left = (double *)malloc(sizeof(double) * NV * nx), *w_right = (double *)malloc(sizeof(double) * NV * nx);
  double *u_left = (double *)malloc(sizeof(double) * NV * nx), *u_right = (double *)malloc(sizeof(double) * NV * nx);
  double *f_left = (double *)malloc(sizeof(double) * NV * nx), *f_right = (double *)malloc(sizeof(double) * NV * nx);
  double *fr = (double *)calloc(NV * nf, sizeof(double)), *s = (double *)malloc(sizeof(double) * NV * nx);
  const double e5 = (double)1e-5f;
  for (int i = 1; i <= nx; ++i) x[i - 1] = (double)((float)i - 0.5f) * dx;
  for (int i = 1; i <= nf; ++i) xf[i - 1] = (double)(i - 1) * dx;
  for (int i = 0; i < nx; ++i) b1_prim(p, u + NV * i, w + NV * i);
  for (int i = 0; i < nx; ++i) {
    const double *wi = w + NV * i;
    double phi_c = 1.0 * x[i], phi_l = 1.0 * xf[i], phi_r = 1.0 * xf[i + 1];
    double h = fmax(wi[2], e5) / fmax(e5, wi[0]) * (1 + (double)1.f / (gamma - 1));
    double h0_left = h + phi_c - phi_l, h0_right = h + phi_c - phi_r;
    double Kapp = fmax(e5, wi[2]) / pow(fmax(e5, wi[0]), gamma);
    double *wl = w_left + NV * i, *wr = w_right + NV * i;
    wl[1] = wi[1]; wr[1] = wi[1];
    wl[0] = pow(((double)1.f / Kapp) * (gamma - 1) / gamma * h0_left, (1 / (gamma - 1)));
    wl[2] = pow(((double)1.f / Kapp), (1 / (gamma - 1))) * pow((gamma - 1) / gamma * h0_left, (gamma / (gamma - 1)));
    wr[0] = pow(((double)1.f / Kapp) * (gamma - 1) / gamma * h0_right, (1 / (gamma - 1)));
    wr[2] = pow(((double)1.f / Kapp), (1 / (gamma - 1))) * pow((gamma - 1) / gamma * h0_right, (gamma / (gamma - 1)));
    b1_cons(p, wl, u_left + NV * i);
    b1_cons(p, wr, u_right + NV * i);
    b1_flux(p, u_left + NV * i, f_left + NV * i);
    b1_flux(p, u_right + NV * i, f_right + NV * i);
  }
  for (int iface = 2; iface <= nf - 1; ++iface) {          /* faces 1 and nx+1 are never computed (:648) */
    int il = iface - 1, ir = iface;
    b1_llflux(p, u_right + NV * (il - 1), u_left + NV * (ir - 1), f_right + NV * (il - 1), f_left + NV * (ir - 1),
              fr + NV * (iface - 1));
  }
  {   /* get_source_rg :409-433 */
    double delta = (double)1.f / (double)nx;
    for (int i = 0; i < nx; ++i) {
      double xm = (i == 0) ? x[0] - delta : x[i - 1];
      double xp = (i == nx - 1) ? x[nx - 1] + delta : x[i + 1];
      s[NV * i + 0] = 0;
      s[NV * i + 1] = (w_right[NV * i + 2] - w_left[NV * i + 2]) / delta;
      s[NV * i + 2] = -w[NV * i + 0] * w[NV * i + 1] * 1 * (xp - xm) / (2 * delta);
    }
  }
  for (int i = 0; i < nx; ++i)
    for (int v = 0; v < NV; ++v) dudt[NV * i + v] = -(fr[NV * (i + 1) + v] - fr[NV * i + v]) * oneoverdx + s[NV * i + v];
  for (int v = 0; v < NV; ++v) { dudt[v] = dudt[NV + v]; dudt[NV * (nx - 1) + v] = dudt[NV * (nx - 2) + v]; }
  free(x); free(xf); free(w); free(w_left); free(w_right); free(u_left); free(u_right); free(f_left); free(f_right);
  free(fr); free(s);
}

/* :200-261 evolve.  'FVM' and 'EQL' evaluate the second stage at u, not w1 (:228, :236) -- reproduced. */
void orc_fv1d_evolve(const orc_fv1d_params *p, double *u, const double *w_eq, double tend, int max_iter, int *iters,
                     double *t_out, double *dt_out) {
  const int n = NV * p->nx;
  const double dx = p->boxlen / (double)p->nx;
  double *dudt = (double *)malloc(sizeof(double) * n), *w1 = (double *)malloc(sizeof(double) * n);
  double t = 0, dt = 0, cmax;
  int iter = 0;
  while (t < tend && (max_iter < 0 || iter < max_iter)) {
    orc_fv1d_compute_max_speed(p, u, &cmax);
    dt = (double)0.8f * dx / cmax / (2.0 * (double)1 + 1.0);
    if (p->solver == 1) {
      orc_fv1d_compute_update_fvm(p, u, w_eq, dudt);
      for (int k = 0; k < n; ++k) w1[k] = u[k] + dt * dudt[k];
      orc_fv1d_compute_update_fvm(p, u, w_eq, dudt);
      for (int k = 0; k < n; ++k) u[k] = 0.5 * u[k] + 0.5 * w1[k] + 0.5 * dt * dudt[k];
    }
    if (p->solver == 2) {
      orc_fv1d_compute_update(p, u, w_eq, dudt);
      for (int k = 0; k < n; ++k) w1[k] = u[k] + dt * dudt[k];
      orc_fv1d_compute_update(p, u, w_eq, dudt);
      for (int k = 0; k < n; ++k) u[k] = 0.5 * u[k] + 0.5 * w1[k] + 0.5 * dt * dudt[k];
    }
    if (p->solver == 3) {
      orc_fv1d_compute_update_sr(p, u, w_eq, dudt);
      for (int k = 0; k < n; ++k) w1[k] = u[k] + dt * dudt[k];
      orc_fv1d_compute_update_sr(p, w1, w_eq, dudt);
      for (int k = 0; k < n; ++k) u[k] = 0.5 * u[k] + 0.5 * w1[k] + 0.5 * dt * dudt[k];
    }
    t = t + dt;
    iter = iter + 1;
  }
  if (iters) *iters = iter;
  if (t_out) *t_out = t;
  if (dt_out) *dt_out = dt;
  free(dudt); free(w1);
}
