/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement (plain C, FP64, no FMA contraction: build with -ffp-contract=off) of the
 * 2D well-balanced first-order finite-volume path of the reference, benchmark_2d.f90.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this file's shared object.
 *
 * PARITY PINNED TO THE REFERENCE'S OWN SOURCE TEXT: the image has no Fortran compiler (so no oracle/_ref build) and
 * the reference ships no golden vectors, but benchmark_2d.f90 itself is EXECUTED, unmodified, by the Fortran-90
 * interpreter oracle/f90interp.py; the vectors it produces (tests/golden/ref_fv2d.npz, generator
 * tests/golden/make_ref_golden.py) are reproduced by this file BIT FOR BIT (tests/test_reference_pins.py:
 * coordinates, initial conditions, equilibria, primitive/conservative, max speed, compute_update_exact,
 * compute_update, whole evolve runs).  Additional pins: dudt == 0 bitwise at the discrete hydrostatic state
 *  and an independent numpy restatement (tests/test_oracle_fv2d.py).
 *
 * Conventions reproduced from the reference (all citations relative to /root/reference):
 *   - arrays are Fortran u(nvar,nx,ny): C offset ((j*nx)+i)*4+v with 0-based i,j,v;
 *   - un-suffixed real literals are real(4) promoted to real(8) (Makefile:3 has empty FFLAGS):
 *     1.21 -> (double)1.21f, 0.3 -> (double)0.3f, (i-0.5) is evaluated in single precision;
 *   - expressions are evaluated left to right, x**2 -> x*x;
 *   - the whole-array temporaries of compute_update_exact are kept (this file is also the
 *     "port" CPU baseline, so it must cost what the reference's structure costs).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NV 4

typedef struct {
  int nx, ny;
  int nequilibrium;     /* parameters_2d.f90:14 */
  double gamma;         /* parameters_2d.f90:19  (1.4 as real(4)) */
  double boxlen_x;      /* parameters_2d.f90:17 */
  double boxlen_y;      /* parameters_2d.f90:18 */
  double cfl;           /* parameters_2d.f90:20 */
} orc_fv2d_params;

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* benchmark_2d.f90:25-43 get_coords.  (i-0.5) is integer minus real(4). */
void orc_fv2d_get_coords(int sx, int sy, double boxlen_x, double boxlen_y, double *x, double *y) {
  double dx = boxlen_x / (double)sx;
  double dy = boxlen_y / (double)sy;
  for (int j = 0; j < sy; ++j)
    for (int i = 0; i < sx; ++i) {
      x[j * sx + i] = (double)((float)(i + 1) - 0.5f) * dx;
      y[j * sx + i] = (double)((float)(j + 1) - 0.5f) * dy;
    }
}

/* benchmark_2d.f90:145-157 compute_primitive (elementwise over n states) */
void orc_fv2d_compute_primitive(const double *u, double *w, double gamma, long n) {
  const double gm1 = gamma - (double)1.0f;
#pragma omp parallel for schedule(static)
  for (long k = 0; k < n; ++k) {
    const double *uu = u + NV * k;
    double *ww = w + NV * k;
    ww[0] = uu[0];
    ww[1] = uu[1] / ww[0];
    ww[2] = uu[2] / ww[0];
    ww[3] = gm1 * (uu[3] - 0.5 * ww[0] * (ww[1] * ww[1] + ww[2] * ww[2]));
  }
}

/* benchmark_2d.f90:159-171 compute_conservative */
void orc_fv2d_compute_conservative(const double *ww, double *u, double gamma, long n) {
  const double gm1 = gamma - (double)1.f;
#pragma omp parallel for schedule(static)
  for (long k = 0; k < n; ++k) {
    const double *w = ww + NV * k;
    double *uu = u + NV * k;
    uu[0] = w[0];
    uu[1] = w[0] * w[1];
    uu[2] = w[0] * w[2];
    uu[3] = w[3] / gm1 + 0.5 * (w[0] * (w[1] * w[1] + w[2] * w[2]));
  }
}

/* benchmark_2d.f90:174-218 get_equilibrium_solution: primitives at (x,y); sx*sy points */
void orc_fv2d_get_equilibrium_solution(int nequilibrium, const double *x, const double *y, double *w,
                                       long n) {
#pragma omp parallel for schedule(static)
  for (long k = 0; k < n; ++k) {
    double *ww = w + NV * k;
    switch (nequilibrium) {
      case 1: {
        double e = exp(-(x[k] + y[k]));
        ww[0] = e; ww[1] = 0.0; ww[2] = 0.0; ww[3] = e;
      } break;
      case 2:
      case 3: {
        double rho_0 = (double)1.21f, p_0 = 1.0, g = 1.0;
        double e = exp(-(rho_0 * g / p_0) * (x[k] + y[k]));
        ww[0] = rho_0 * e; ww[1] = 0.0; ww[2] = 0.0; ww[3] = p_0 * e;
      } break;
      default:  /* case 4 */
        ww[0] = 0.0; ww[1] = 0.0; ww[2] = 0.0; ww[3] = 0.0;
        break;
    }
  }
}

/* benchmark_2d.f90:45-113 get_initial_conditions (returns conservative u) */
void orc_fv2d_get_initial_conditions(int ninit, double eta, double gamma, const double *x,
                                     const double *y, double *u, long n) {
  double *w = (double *)malloc(sizeof(double) * NV * n);
  for (long k = 0; k < n; ++k) {
    double *ww = w + NV * k;
    double xx = x[k], yy = y[k];
    switch (ninit) {
      case 1: {
        double e = exp(-(xx + yy));
        ww[0] = e; ww[1] = 0; ww[2] = 0; ww[3] = e;
      } break;
      case 2: {
        double rho_0 = (double)1.21f, p_0 = 1.0, g = 1.0;
        double e = exp(-(rho_0 * g / p_0) * (xx + yy));
        ww[0] = rho_0 * e; ww[1] = 0; ww[2] = 0; ww[3] = p_0 * e;
      } break;
      case 3: {
        double rho_0 = (double)1.21f, p_0 = 1.0, g = 1.0;
        double e = exp(-(rho_0 * g / p_0) * (xx + yy));
        double dxp = xx - (double)0.3f, dyp = yy - (double)0.3f;
        double bump = exp(-(100.0 * (rho_0 * g / p_0) * (dxp * dxp + dyp * dyp)));
        ww[0] = rho_0 * e; ww[1] = 0; ww[2] = 0; ww[3] = p_0 * e + eta * bump;
      } break;
      default: { /* case 4: 2D Riemann problem */
        if (xx >= 0.5 && yy >= 0.5) {
          ww[0] = 1.5; ww[1] = 0.; ww[2] = 0.; ww[3] = 1.5;
        } else if (xx < 0.5 && yy >= 0.5) {
          ww[0] = (double)0.5323f; ww[1] = (double)1.206f; ww[2] = 0.; ww[3] = (double)0.3f;
        } else if (xx < 0.5 && yy < 0.5) {
          ww[0] = (double)0.138f; ww[1] = (double)1.206f; ww[2] = (double)1.206f; ww[3] = (double)0.029f;
        } else {
          ww[0] = (double)0.5323f; ww[1] = 0.; ww[2] = (double)1.206f; ww[3] = (double)0.3f;
        }
      } break;
    }
  }
  orc_fv2d_compute_conservative(w, u, gamma, n);
  free(w);
}

/* benchmark_2d.f90:283-295 compute_speed */
static inline double speed1(const double *u, double gamma) {
  const double gm1 = gamma - (double)1.0f;
  double w1 = u[0];
  double w2 = u[1] / w1;
  double w3 = u[2] / w1;
  double w4 = gm1 * (u[3] - 0.5 * w1 * (w2 * w2 + w3 * w3));
  double cs = sqrt(gamma * fmax(w4, 1e-10) / fmax(w1, 1e-10));
  return sqrt(w2 * w2 + w3 * w3) + cs;
}

/* benchmark_2d.f90:264-279 compute_max_speed (all cells, boundary included; plain max from 0.0) */
void orc_fv2d_compute_max_speed(const orc_fv2d_params *p, const double *u, double *cmax_out) {
  long n = (long)p->nx * p->ny;
  double cmax = 0.0;
#pragma omp parallel for schedule(static) reduction(max : cmax)
  for (long k = 0; k < n; ++k) {
    double s = speed1(u + NV * k, p->gamma);
    cmax = (s > cmax) ? s : cmax;   /* MAX(cmax,speed) */
  }
  *cmax_out = cmax;
}

/* benchmark_2d.f90:299-325 compute_flux: both directional fluxes of n states.
 * flux layout: [dir][k][v]  (Fortran flux(nvar,sx,sy,2)) */
static void compute_flux(const double *u, double *flux, double gamma, long n) {
  double *w = (double *)malloc(sizeof(double) * NV * n);
  orc_fv2d_compute_primitive(u, w, gamma, n);
  double *f1 = flux, *f2 = flux + NV * n;
#pragma omp parallel for schedule(static)
  for (long k = 0; k < n; ++k) {
    const double *uu = u + NV * k, *ww = w + NV * k;
    double *a = f1 + NV * k, *b = f2 + NV * k;
    a[0] = ww[1] * uu[0];
    a[1] = ww[1] * uu[1] + ww[3];
    a[2] = ww[0] * ww[1] * ww[2];
    a[3] = ww[1] * uu[3] + ww[1] * ww[3];
    b[0] = uu[0] * ww[2];
    b[1] = uu[1] * ww[2];
    b[2] = uu[2] * ww[2] + ww[3];
    b[3] = ww[2] * uu[3] + ww[2] * ww[3];
  }
  free(w);
}

/* benchmark_2d.f90:327-350 get_source (phi_x = phi_y = 1.) */
static void get_source(const double *w, double *s, long n) {
  const double phi_x = 1.0, phi_y = 1.0;
#pragma omp parallel for schedule(static)
  for (long k = 0; k < n; ++k) {
    const double *ww = w + NV * k;
    double *ss = s + NV * k;
    ss[0] = 0.0;
    ss[1] = -ww[0] * phi_x;
    ss[2] = -ww[0] * phi_y;
    ss[3] = -ww[0] * (ww[1] * phi_x + ww[2] * phi_y);
  }
}

/* benchmark_2d.f90:353-367 compute_llflux */
static inline void compute_llflux(const double *uleft, const double *uright, const double *f_left,
                                  const double *f_right, double *fgdnv, double gamma) {
  double cleft = speed1(uleft, gamma);
  double cright = speed1(uright, gamma);
  double cmax = (cleft > cright) ? cleft : cright;
  for (int v = 0; v < NV; ++v)
    fgdnv[v] = 0.5 * (f_right[v] + f_left[v]) + 0.5 * cmax * (uleft[v] - uright[v]);
}

#define U3(a, i, j, nxx) ((a) + NV * ((long)(j) * (nxx) + (i)))

/* benchmark_2d.f90:465-618 compute_update_exact(u, w_eq, dudt).
 * The debug `write(*,*) dudt` at :610 is omitted (no arithmetic effect). */
void orc_fv2d_compute_update_exact(const orc_fv2d_params *p, const double *u, const double *w_eq,
                                   double *dudt) {
  const int nx = p->nx, ny = p->ny;
  const long n = (long)nx * ny;
  const int nxf = nx + 1, nyf = ny + 1;
  const long nf = (long)nxf * nyf;
  const double gamma = p->gamma;
  const double dx = p->boxlen_x / (double)nx;
  const double dy = p->boxlen_y / (double)ny;
  const double oneoverdx = 1 / dx;
  const double oneoverdy = 1 / dy;

  double *u_eq = (double *)malloc(sizeof(double) * NV * n);
  double *delta_u = (double *)malloc(sizeof(double) * NV * n);
  double *u_left = (double *)malloc(sizeof(double) * NV * n);
  double *u_right = (double *)malloc(sizeof(double) * NV * n);
  double *u_top = (double *)malloc(sizeof(double) * NV * n);
  double *u_bottom = (double *)malloc(sizeof(double) * NV * n);
  double *flux_left = (double *)malloc(sizeof(double) * 2 * NV * n);
  double *flux_right = (double *)malloc(sizeof(double) * 2 * NV * n);
  double *flux_top = (double *)malloc(sizeof(double) * 2 * NV * n);
  double *flux_bottom = (double *)malloc(sizeof(double) * 2 * NV * n);
  double *F = (double *)calloc(NV * (long)nxf * ny, sizeof(double));
  double *G = (double *)calloc(NV * (long)nx * nyf, sizeof(double));
  double *F_eq = (double *)malloc(sizeof(double) * 2 * NV * nf);
  double *G_eq = (double *)malloc(sizeof(double) * 2 * NV * nf);
  double *u_x_faces = (double *)malloc(sizeof(double) * NV * nf);
  double *w_x_faces = (double *)malloc(sizeof(double) * NV * nf);
  double *u_y_faces = (double *)malloc(sizeof(double) * NV * nf);
  double *w_y_faces = (double *)malloc(sizeof(double) * NV * nf);
  double *x_faces = (double *)malloc(sizeof(double) * nf);
  double *y_faces = (double *)malloc(sizeof(double) * nf);
  double *x = (double *)malloc(sizeof(double) * nf);
  double *y = (double *)malloc(sizeof(double) * nf);
  double *w = (double *)malloc(sizeof(double) * NV * n);
  double *s = (double *)malloc(sizeof(double) * NV * n);
  double *s_eq = (double *)malloc(sizeof(double) * NV * n);

  /* :496 */
  orc_fv2d_compute_conservative(w_eq, u_eq, gamma, n);
  /* :499 */
#pragma omp parallel for schedule(static)
  for (long k = 0; k < NV * n; ++k) delta_u[k] = u[k] - u_eq[k];

  /* :505-522 face / centre coordinates on the (nx+1,ny+1) arrays; y_faces uses dx (sic, :513) */
#pragma omp parallel for schedule(static)
  for (int j = 0; j < nyf; ++j)
    for (int i = 0; i < nxf; ++i) {
      long k = (long)j * nxf + i;
      x_faces[k] = (double)(i) * dx;                       /* (i-1)*dx, 1-based i */
      y_faces[k] = (double)(j) * dx;                       /* (j-1)*dx  (sic) */
      x[k] = (double)((float)(i + 1) - 0.5f) * dx;
      y[k] = (double)((float)(j + 1) - 0.5f) * dy;
    }

  /* :524-527 */
  orc_fv2d_get_equilibrium_solution(p->nequilibrium, x_faces, y, w_x_faces, nf);
  orc_fv2d_get_equilibrium_solution(p->nequilibrium, x, y_faces, w_y_faces, nf);
  orc_fv2d_compute_conservative(w_x_faces, u_x_faces, gamma, nf);
  orc_fv2d_compute_conservative(w_y_faces, u_y_faces, gamma, nf);

  /* :533-537 */
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i)
      for (int v = 0; v < NV; ++v) {
        double d = U3(delta_u, i, j, nx)[v];
        U3(u_left, i, j, nx)[v] = U3(u_x_faces, i, j, nxf)[v] + d;
        U3(u_right, i, j, nx)[v] = U3(u_x_faces, i + 1, j, nxf)[v] + d;
        U3(u_top, i, j, nx)[v] = U3(u_y_faces, i, j + 1, nxf)[v] + d;
        U3(u_bottom, i, j, nx)[v] = U3(u_y_faces, i, j, nxf)[v] + d;
      }

  /* :546-549 */
  compute_flux(u_left, flux_left, gamma, n);
  compute_flux(u_right, flux_right, gamma, n);
  compute_flux(u_top, flux_top, gamma, n);
  compute_flux(u_bottom, flux_bottom, gamma, n);

  /* :552-567 x sweep: F(1:nvar, iface, j), iface = 1..nx+1 */
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; ++j)
    for (int iface = 0; iface <= nx; ++iface) {
      int ileft = iface - 1, iright = iface;
      if (iface == 0) ileft = 0;
      if (iface == nx) iright = nx - 1;
      compute_llflux(U3(u_right, ileft, j, nx), U3(u_left, iright, j, nx),
                     U3(flux_right, ileft, j, nx), U3(flux_left, iright, j, nx),
                     U3(F, iface, j, nxf), gamma);
    }
  /* :570-584 y sweep: G(1:nvar, i, jface), jface = 1..ny+1, uses direction-2 fluxes */
#pragma omp parallel for schedule(static)
  for (int jface = 0; jface <= ny; ++jface)
    for (int i = 0; i < nx; ++i) {
      int ileft = jface - 1, iright = jface;
      if (jface == 0) ileft = 0;
      if (jface == ny) iright = ny - 1;
      compute_llflux(U3(u_top, i, ileft, nx), U3(u_bottom, i, iright, nx),
                     U3(flux_top + NV * n, i, ileft, nx), U3(flux_bottom + NV * n, i, iright, nx),
                     U3(G, i, jface, nx), gamma);
    }

  /* :588-590 */
  get_source(w_eq, s_eq, n);
  orc_fv2d_compute_primitive(u, w, gamma, n);
  get_source(w, s, n);

  /* :596-597 */
  compute_flux(u_x_faces, F_eq, gamma, nf);
  compute_flux(u_y_faces, G_eq, gamma, nf);
  const double *F_eq1 = F_eq;            /* F_eq(:,:,:,1) */
  const double *G_eq2 = G_eq + NV * nf;  /* G_eq(:,:,:,2) */

  /* :599-609 */
#pragma omp parallel for schedule(static)
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i)
      for (int v = 0; v < NV; ++v) {
        double r = -(U3(F, i + 1, j, nxf)[v] - U3(F, i, j, nxf)[v]) * oneoverdx
                   - (U3(G, i, j + 1, nx)[v] - U3(G, i, j, nx)[v]) * oneoverdy;
        r = r + U3(s, i, j, nx)[v];
        r = r - U3(s_eq, i, j, nx)[v];
        r = r + (U3(F_eq1, i + 1, j, nxf)[v] - U3(F_eq1, i, j, nxf)[v]) * oneoverdx;
        r = r + (U3(G_eq2, i, j + 1, nxf)[v] - U3(G_eq2, i, j, nxf)[v]) * oneoverdy;
        U3(dudt, i, j, nx)[v] = r;
      }
  /* :611-614 */
  for (int j = 0; j < ny; ++j)
    for (int v = 0; v < NV; ++v) {
      U3(dudt, 0, j, nx)[v] = 0.0;
      U3(dudt, nx - 1, j, nx)[v] = 0.0;
    }
  for (int i = 0; i < nx; ++i)
    for (int v = 0; v < NV; ++v) {
      U3(dudt, i, 0, nx)[v] = 0.0;
      U3(dudt, i, ny - 1, nx)[v] = 0.0;
    }

  free(u_eq); free(delta_u); free(u_left); free(u_right); free(u_top); free(u_bottom);
  free(flux_left); free(flux_right); free(flux_top); free(flux_bottom); free(F); free(G);
  free(F_eq); free(G_eq); free(u_x_faces); free(w_x_faces); free(u_y_faces); free(w_y_faces);
  free(x_faces); free(y_faces); free(x); free(y); free(w); free(s); free(s_eq);
}

/* benchmark_2d.f90:370-463 compute_update (plain, non well-balanced; not called by evolve).
 * Reproduces `iright = ny` at x-face nx+1 (:418): only feeds the zeroed cell i=nx. */
void orc_fv2d_compute_update(const orc_fv2d_params *p, const double *u, const double *w_eq,
                             double *dudt) {
  (void)w_eq;
  const int nx = p->nx, ny = p->ny;
  const long n = (long)nx * ny;
  const int nxf = nx + 1, nyf = ny + 1;
  const double gamma = p->gamma;
  const double dx = p->boxlen_x / (double)nx;
  const double dy = p->boxlen_y / (double)ny;
  const double oneoverdx = 1 / dx, oneoverdy = 1 / dy;
  double *flux = (double *)malloc(sizeof(double) * 2 * NV * n);
  double *F = (double *)calloc(NV * (long)nxf * ny, sizeof(double));
  double *G = (double *)calloc(NV * (long)nx * nyf, sizeof(double));
  double *w = (double *)malloc(sizeof(double) * NV * n);
  double *s = (double *)malloc(sizeof(double) * NV * n);
  /* u_left = u_right = u_top = u_bottom = u, so the four compute_flux calls give one result */
  compute_flux(u, flux, gamma, n);
  for (int j = 0; j < ny; ++j)
    for (int iface = 0; iface <= nx; ++iface) {
      int ileft = iface - 1, iright = iface;
      if (iface == 0) ileft = 0;
      if (iface == nx) iright = ny - 1; /* sic */
      if (iright > nx - 1) iright = nx - 1; /* keep the restatement in bounds when ny > nx */
      compute_llflux(U3(u, ileft, j, nx), U3(u, iright, j, nx), U3(flux, ileft, j, nx),
                     U3(flux, iright, j, nx), U3(F, iface, j, nxf), gamma);
    }
  for (int jface = 0; jface <= ny; ++jface)
    for (int i = 0; i < nx; ++i) {
      int ileft = jface - 1, iright = jface;
      if (jface == 0) ileft = 0;
      if (jface == ny) iright = ny - 1;
      compute_llflux(U3(u, i, ileft, nx), U3(u, i, iright, nx), U3(flux + NV * n, i, ileft, nx),
                     U3(flux + NV * n, i, iright, nx), U3(G, i, jface, nx), gamma);
    }
  orc_fv2d_compute_primitive(u, w, gamma, n);
  get_source(w, s, n);
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i)
      for (int v = 0; v < NV; ++v)
        U3(dudt, i, j, nx)[v] = -(U3(F, i + 1, j, nxf)[v] - U3(F, i, j, nxf)[v]) * oneoverdx
                                - (U3(G, i, j + 1, nx)[v] - U3(G, i, j, nx)[v]) * oneoverdy
                                + U3(s, i, j, nx)[v];
  for (int j = 0; j < ny; ++j)
    for (int v = 0; v < NV; ++v) { U3(dudt, 0, j, nx)[v] = 0.0; U3(dudt, nx - 1, j, nx)[v] = 0.0; }
  for (int i = 0; i < nx; ++i)
    for (int v = 0; v < NV; ++v) { U3(dudt, i, 0, nx)[v] = 0.0; U3(dudt, i, ny - 1, nx)[v] = 0.0; }
  free(flux); free(F); free(G); free(w); free(s);
}

/* benchmark_2d.f90:221-260 evolve(u,u_eq): SSP-RK2 with dt = 0.5*dx/cmax*cfl, no clamp to tend.
 * max_iter < 0 means "until t >= tend" as in the reference. The `write(*,*)'time='` line is omitted. */
void orc_fv2d_evolve(const orc_fv2d_params *p, double *u, const double *w_eq, double tend,
                     int max_iter, int *iters_out, double *t_out, double *dt_out, double *cmax_out) {
  const long n = (long)p->nx * p->ny;
  const double dx = p->boxlen_x / (double)p->nx;
  double *dudt = (double *)malloc(sizeof(double) * NV * n);
  double *w1 = (double *)malloc(sizeof(double) * NV * n);
  double t = 0, dt = 0, cmax = 0;
  int iter = 0;
  while (t < tend && (max_iter < 0 || iter < max_iter)) {
    orc_fv2d_compute_max_speed(p, u, &cmax);
    dt = 0.5 * dx / cmax * p->cfl;
    orc_fv2d_compute_update_exact(p, u, w_eq, dudt);
#pragma omp parallel for schedule(static)
    for (long k = 0; k < NV * n; ++k) w1[k] = u[k] + dt * dudt[k];
    orc_fv2d_compute_update_exact(p, w1, w_eq, dudt);
#pragma omp parallel for schedule(static)
    for (long k = 0; k < NV * n; ++k) u[k] = 0.5 * u[k] + 0.5 * w1[k] + 0.5 * dt * dudt[k];
    t = t + dt;
    iter = iter + 1;
  }
  if (iters_out) *iters_out = iter;
  if (t_out) *t_out = t;
  if (dt_out) *dt_out = dt;
  if (cmax_out) *cmax_out = cmax;
  free(dudt); free(w1);
}
