// 1D finite-volume paths: fvm.f90 (plain FV + centred gravity source) and benchmark_1d.f90 ('FVM', 'EQL'
// equilibrium subtraction, 'WB1' local hydrostatic reconstruction), SSP-RK2 time loop on the device.
//
// These are the reference's correctness configurations (nx ~ 128-512): one thread per cell, every face flux the
// cell needs recomputed in the reference's operation order (file compiled with -fmad=false), data kept in the
// reference's own u(nvar,nx) layout.  Differences to the CPU restatement come only from exp()/pow().
#include "common.cuh"
#include <algorithm>
#include <cmath>

namespace wb { namespace fv1d {

constexpr int NV = 3;

struct Ctrl1 {
  double t, dt, tend, cmax;
  int iter, max_iter, skip;
};

struct P1 {            // both programs
  int nx, bc, source, nequilibrium, solver;
  double gamma, boxlen, dt_den;   // dt = 0.8*dx/cmax/dt_den
};

// compute_primitive (fvm.f90:176-186, benchmark_1d.f90:103-112)
__device__ __forceinline__ void prim(const P1& p, const double* u, double* w) {
  w[0] = u[0];
  w[1] = u[1] / w[0];
  w[2] = (p.gamma - (double)1.0f) * (u[2] - 0.5 * w[0] * (w[1] * w[1]));
}
__device__ __forceinline__ void cons(const P1& p, const double* w, double* u) {
  u[0] = w[0];
  u[1] = w[0] * w[1];
  u[2] = w[2] / (p.gamma - (double)1.0f) + 0.5 * w[0] * (w[1] * w[1]);
}
__device__ __forceinline__ double speed(const P1& p, const double* u) {
  double w[NV];
  prim(p, u, w);
  double cs = sqrt(p.gamma * fmax(w[2], 1e-10) / fmax(w[0], 1e-10));
  return fabs(w[1]) + cs;
}
__device__ __forceinline__ void flux(const P1& p, const double* u, double* f) {
  double w[NV];
  prim(p, u, w);
  f[0] = w[1] * u[0];
  f[1] = w[1] * u[1] + w[2];
  f[2] = w[1] * u[2] + w[2] * w[1];
}
// compute_llflux with precomputed physical fluxes (benchmark_1d.f90:437-451; fvm.f90:270-283 computes them inside)
__device__ __forceinline__ void llflux(const P1& p, const double* ul, const double* ur, const double* fl, const double* fr,
                                       double* fg) {
  double cmax = fmax(speed(p, ul), speed(p, ur));
#pragma unroll
  for (int v = 0; v < NV; ++v) fg[v] = 0.5 * (fr[v] + fl[v]) - 0.5 * cmax * (ur[v] - ul[v]);
}
__device__ __forceinline__ void llflux_u(const P1& p, const double* ul, const double* ur, double* fg) {
  double fl[NV], fr[NV];
  flux(p, ul, fl);
  flux(p, ur, fr);
  llflux(p, ul, ur, fl, fr, fg);
}

// ------------------------------------------------------------------------------------ fvm.f90:188-264
__global__ void k_fvm_update(const double* __restrict__ u, double* __restrict__ dudt, P1 p, const Ctrl1* ctrl) {
  if (ctrl && ctrl->skip) return;
  int ic = blockIdx.x * blockDim.x + threadIdx.x + 1;     // 1-based cell
  if (ic > p.nx) return;
  const int nx = p.nx;
  const double dx = p.boxlen / (double)nx, oneoverdx = 1.0 / dx;
  int um = ic - 1, up = ic + 1;
  if (p.bc == 1) { if (ic == 1) um = nx; if (ic == nx) up = 1; }
  if (p.bc == 2) { if (ic == 1) um = 1; if (ic == nx) up = nx; }
  double fl[NV], fr[NV], s[NV] = {0.0, 0.0, 0.0};
  llflux_u(p, u + NV * (um - 1), u + NV * (ic - 1), fl);
  llflux_u(p, u + NV * (ic - 1), u + NV * (up - 1), fr);
  if (p.source == 2) {
    double x_minus = ((double)(ic - 1) - 0.5) * dx, x_plus = ((double)(ic + 1) - 0.5) * dx;
    if (ic == 1) x_minus = ((double)1 - 0.5) * dx;
    if (ic == nx) x_plus = ((double)nx - 0.5) * dx;
    double w[NV];
    prim(p, u + NV * (ic - 1), w);
    s[0] = 0;
    s[1] = -w[0] * 1 * (x_plus - x_minus) / (2 * dx);
    s[2] = -w[0] * w[1] * 1 * (x_plus - x_minus) / (2 * dx);
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) dudt[NV * (ic - 1) + v] = -oneoverdx * (fr[v] - fl[v]) + s[v];
}

// ------------------------------------------------------------------------------------ benchmark_1d.f90
// get_equilibrium_solution :125-152
__device__ __forceinline__ void eq_prim(const P1& p, double x, double* w) {
  const double gamma = p.gamma;
  if (p.nequilibrium == 3) {
    double base = (1 - ((gamma - 1) / gamma) * 1 * x);
    w[0] = pow(base, (1 / (gamma - 1)));
    w[1] = 0;
    w[2] = pow(base, (gamma / (gamma - 1)));
  } else {
    w[0] = exp(-x); w[1] = 0; w[2] = exp(-x);
  }
}
__device__ __forceinline__ double xc(int i, double dx) { return (double)((float)i - 0.5f) * dx; }   // 1-based
__device__ __forceinline__ double xf(int i, double dx) { return (double)(i - 1) * dx; }

__device__ __forceinline__ void face_indices(const P1& p, int iface, int& il, int& ir) {
  il = iface - 1; ir = iface;
  if (p.bc == 1) { if (iface == 1) il = p.nx; if (iface == p.nx + 1) ir = 1; }
  if (p.bc == 2 || p.bc == 3) { if (iface == 1) il = 1; if (iface == p.nx + 1) ir = p.nx; }
}
// get_source :380-406 for cell i (callers pass CONSERVATIVE variables as `w`, :365-366, :541)
__device__ __forceinline__ void get_source(const P1& p, const double* w, int i, double* s) {
  const int nx = p.nx;
  const double dx = p.boxlen / (double)nx, delta = 1 / (double)nx;
  double xm = (i == 1) ? xc(1, dx) - delta : xc(i - 1, dx);
  double xp = (i == nx) ? xc(nx, dx) + delta : xc(i + 1, dx);
  s[0] = 0;
  s[1] = -w[0] * 1 * (xp - xm) / (2 * delta);
  s[2] = -w[0] * w[1] * 1 * (xp - xm) / (2 * delta);
}

// SCHEME 1 'FVM' (:454-549), 2 'EQL' (:263-377).  Thread i evaluates the cell ie = clamp(i, 2, nx-1)
// (dudt(:,1) = dudt(:,2), dudt(:,nx) = dudt(:,nx-1)).
template <int SCHEME>
__global__ void k_b1_update(const double* __restrict__ u, const double* __restrict__ w_eq, double* __restrict__ dudt, P1 p,
                            const Ctrl1* ctrl) {
  if (ctrl && ctrl->skip) return;
  int it = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (it > p.nx) return;
  const int nx = p.nx;
  const int i = min(max(it, 2), nx - 1);
  const double dx = p.boxlen / (double)nx, oneoverdx = 1 / dx;
  auto delta_w = [&](int c, double* d) {
    double ue[NV];
    cons(p, w_eq + NV * (c - 1), ue);
#pragma unroll
    for (int v = 0; v < NV; ++v) d[v] = u[NV * (c - 1) + v] - ue[v];
  };
  auto ueq_face = [&](int f, double* uf) {
    double w[NV];
    eq_prim(p, xf(f, dx), w);
    cons(p, w, uf);
  };
  // state on the low/high side of face `iface` and its physical flux
  auto side = [&](int c, int face_of_c /*0 = left face value u_left(c), 1 = right face value u_right(c)*/, double* us) {
    if (SCHEME == 1) {
#pragma unroll
      for (int v = 0; v < NV; ++v) us[v] = u[NV * (c - 1) + v];
    } else {
      double d[NV], uf[NV];
      delta_w(c, d);
      ueq_face(c + face_of_c, uf);
#pragma unroll
      for (int v = 0; v < NV; ++v) us[v] = d[v] + uf[v];
    }
  };
  auto riemann = [&](int iface, double* fr) {
    int il, ir;
    face_indices(p, iface, il, ir);
    double ul[NV], ur[NV], fl[NV], frr[NV];
    side(il, 1, ul);
    side(ir, 0, ur);
    flux(p, ul, fl);
    flux(p, ur, frr);
    llflux(p, ul, ur, fl, frr, fr);
    if (p.bc == 3 && iface == 1) {
      double w_minus[NV] = {1., 0., 1.}, u_face[NV], d[NV], a[NV], f_minus[NV];
      cons(p, w_minus, u_face);
      delta_w(1, d);
#pragma unroll
      for (int v = 0; v < NV; ++v) a[v] = u_face[v] + d[v];
      flux(p, a, f_minus);
      llflux(p, a, ur, f_minus, frr, fr);
    }
    if (p.bc == 3 && iface == nx + 1) {
      double w_plus[NV] = {1., 0., 1.}, u_plus[NV], d[NV], a[NV], b[NV], f_plus[NV];
      cons(p, w_plus, u_plus);
      delta_w(nx, d);
#pragma unroll
      for (int v = 0; v < NV; ++v) { a[v] = u_plus[v] + d[v]; b[v] = ul[v] + d[v]; }
      flux(p, a, f_plus);
      llflux(p, b, u_plus, fl, f_plus, fr);
    }
  };
  double f0[NV], f1[NV], s[NV];
  riemann(i, f0);
  riemann(i + 1, f1);
  get_source(p, u + NV * (i - 1), i, s);
  if (SCHEME == 1) {
#pragma unroll
    for (int v = 0; v < NV; ++v) dudt[NV * (it - 1) + v] = -(f1[v] - f0[v]) * oneoverdx + s[v];
  } else {
    double ue[NV], se[NV], uf0[NV], uf1[NV], e0[NV], e1[NV];
    cons(p, w_eq + NV * (i - 1), ue);
    get_source(p, ue, i, se);
    ueq_face(i, uf0); ueq_face(i + 1, uf1);
    flux(p, uf0, e0); flux(p, uf1, e1);
#pragma unroll
    for (int v = 0; v < NV; ++v)
      dudt[NV * (it - 1) + v] = -(f1[v] - f0[v]) * oneoverdx + s[v] + (e1[v] - e0[v]) * oneoverdx - se[v];
  }
}

// 'WB1' compute_update_sr :553-747 (phi(x) = x, :749-755); faces 2..nx only (:648)
__global__ void k_b1_update_sr(const double* __restrict__ u, double* __restrict__ dudt, P1 p, const Ctrl1* ctrl) {
  if (ctrl && ctrl->skip) return;
  int it = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (it > p.nx) return;
  const int nx = p.nx;
  const int i = min(max(it, 2), nx - 1);
  const double gamma = p.gamma;
  const double dx = p.boxlen / (double)nx, oneoverdx = p.boxlen / dx;   // sic :574
  const double e5 = (double)1e-5f;
  // hydrostatic reconstruction of cell c: primitive states at its left and right face
  auto recon = [&](int c, double* wl, double* wr, double* wc) {
    prim(p, u + NV * (c - 1), wc);
    double phi_c = 1.0 * xc(c, dx), phi_l = 1.0 * xf(c, dx), phi_r = 1.0 * xf(c + 1, dx);
    double h = fmax(wc[2], e5) / fmax(e5, wc[0]) * (1 + (double)1.f / (gamma - 1));
    double h0_left = h + phi_c - phi_l, h0_right = h + phi_c - phi_r;
    double Kapp = fmax(e5, wc[2]) / pow(fmax(e5, wc[0]), gamma);
    wl[1] = wc[1]; wr[1] = wc[1];
    wl[0] = pow(((double)1.f / Kapp) * (gamma - 1) / gamma * h0_left, (1 / (gamma - 1)));
    wl[2] = pow(((double)1.f / Kapp), (1 / (gamma - 1))) * pow((gamma - 1) / gamma * h0_left, (gamma / (gamma - 1)));
    wr[0] = pow(((double)1.f / Kapp) * (gamma - 1) / gamma * h0_right, (1 / (gamma - 1)));
    wr[2] = pow(((double)1.f / Kapp), (1 / (gamma - 1))) * pow((gamma - 1) / gamma * h0_right, (gamma / (gamma - 1)));
  };
  double wl[3][NV], wr[3][NV], wc[3][NV];          // cells i-1, i, i+1
  for (int k = 0; k < 3; ++k) recon(i - 1 + k, wl[k], wr[k], wc[k]);
  auto riemann = [&](int lo /*index into the 3 cells of the low side*/, double* fr) {
    double ul[NV], ur[NV], fl[NV], frr[NV];
    cons(p, wr[lo], ul);          // u_right(ileft)
    cons(p, wl[lo + 1], ur);      // u_left(iright)
    flux(p, ul, fl);
    flux(p, ur, frr);
    llflux(p, ul, ur, fl, frr, fr);
  };
  double f0[NV] = {0.0, 0.0, 0.0}, f1[NV] = {0.0, 0.0, 0.0};
  // face i (between i-1 and i) exists for i >= 2; face i+1 for i+1 <= nx  (always true for the clamped i)
  riemann(0, f0);
  riemann(1, f1);
  // get_source_rg :409-433
  const double delta = (double)1.f / (double)nx;
  double xm = (i == 1) ? xc(1, dx) - delta : xc(i - 1, dx);
  double xp = (i == nx) ? xc(nx, dx) + delta : xc(i + 1, dx);
  double s[NV];
  s[0] = 0;
  s[1] = (wr[1][2] - wl[1][2]) / delta;
  s[2] = -wc[1][0] * wc[1][1] * 1 * (xp - xm) / (2 * delta);
#pragma unroll
  for (int v = 0; v < NV; ++v) dudt[NV * (it - 1) + v] = -(f1[v] - f0[v]) * oneoverdx + s[v];
}

// max speed (single block; nx is small), optionally the step's dt and skip flag
__global__ void k1_max_speed(const double* __restrict__ u, P1 p, Ctrl1* ctrl, int set_dt) {
  __shared__ double sh[256];
  double m = 0.0;
  for (int i = threadIdx.x; i < p.nx; i += blockDim.x) m = fmax(m, speed(p, u + NV * i));
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
      if (!done) ctrl->dt = (double)0.8f * (p.boxlen / (double)p.nx) / sh[0] / p.dt_den;
    }
  }
}
// w1 = u + dt*dudt   |   u = 0.5*u + 0.5*w1 + 0.5*dt*dudt
__global__ void k1_axpy(double* out, const double* a, const double* b, const double* d, int n, int stage, const Ctrl1* ctrl) {
  if (ctrl->skip) return;
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double dt = ctrl->dt;
  if (stage == 1) out[k] = a[k] + dt * d[k];
  else out[k] = 0.5 * a[k] + 0.5 * b[k] + 0.5 * dt * d[k];
}
__global__ void k1_advance(Ctrl1* c) {
  if (c->skip) return;
  c->t = c->t + c->dt;
  c->iter = c->iter + 1;
}
__global__ void k1_ctrl_init(Ctrl1* c, double tend, int max_iter, int reset) {
  if (reset) { c->t = 0.0; c->iter = 0; c->dt = 0.0; }
  c->tend = tend; c->max_iter = max_iter; c->skip = 0;
}

struct Handle1 {
  P1 p;
  int dev = 0;
  cudaStream_t stream = nullptr;
  double *u = nullptr, *w1 = nullptr, *dudt = nullptr, *weq = nullptr;
  Ctrl1* ctrl = nullptr;
  Ctrl1* h_ctrl = nullptr;
  bool is_fvm = false;
};

static int h1_create(Handle1** out, const P1& p, int device, bool is_fvm) {
  int dev = 0;
  WB_CHECK(select_device(device, &dev));
  Handle1* h = new Handle1;
  h->p = p; h->dev = dev; h->is_fvm = is_fvm;
  size_t fb = sizeof(double) * NV * p.nx;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&h->u, fb);
  if (e == cudaSuccess) e = cudaMalloc(&h->w1, fb);
  if (e == cudaSuccess) e = cudaMalloc(&h->dudt, fb);
  if (e == cudaSuccess) e = cudaMalloc(&h->weq, fb);
  if (e == cudaSuccess) e = cudaMalloc(&h->ctrl, sizeof(Ctrl1));
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_ctrl, sizeof(Ctrl1));
  if (e == cudaSuccess) e = cudaMemset(h->ctrl, 0, sizeof(Ctrl1));
  if (e == cudaSuccess) e = cudaMemset(h->weq, 0, fb);
  if (e != cudaSuccess) {
    set_error("1D handle allocation failed: %s", cudaGetErrorString(e));
    delete h;
    return WB_ERR_CUDA;
  }
  *out = h;
  return WB_OK;
}
static void h1_destroy(Handle1* h) {
  if (!h) return;
  cudaSetDevice(h->dev);
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  cudaFree(h->u); cudaFree(h->w1); cudaFree(h->dudt); cudaFree(h->weq); cudaFree(h->ctrl);
  if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
  delete h;
}
// RHS of the configured scheme: in -> h->dudt
static int h1_rhs(Handle1* h, const double* in, int scheme, bool use_ctrl) {
  const Ctrl1* c = use_ctrl ? h->ctrl : nullptr;
  dim3 b(128), g((h->p.nx + 127) / 128);
  if (h->is_fvm) k_fvm_update<<<g, b, 0, h->stream>>>(in, h->dudt, h->p, c);
  else if (scheme == 1) k_b1_update<1><<<g, b, 0, h->stream>>>(in, h->weq, h->dudt, h->p, c);
  else if (scheme == 2) k_b1_update<2><<<g, b, 0, h->stream>>>(in, h->weq, h->dudt, h->p, c);
  else k_b1_update_sr<<<g, b, 0, h->stream>>>(in, h->dudt, h->p, c);
  WB_LAUNCH_CHECK();
  return WB_OK;
}
static int h1_update_host(Handle1* h, const double* u, const double* w_eq, double* dudt, int scheme) {
  WB_CUDA(cudaSetDevice(h->dev));
  size_t fb = sizeof(double) * NV * h->p.nx;
  WB_CUDA(cudaMemcpyAsync(h->u, u, fb, cudaMemcpyHostToDevice, h->stream));
  if (w_eq) WB_CUDA(cudaMemcpyAsync(h->weq, w_eq, fb, cudaMemcpyHostToDevice, h->stream));
  WB_CHECK(h1_rhs(h, h->u, scheme, false));
  WB_CUDA(cudaMemcpyAsync(dudt, h->dudt, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  return WB_OK;
}
static int h1_max_speed_host(Handle1* h, const double* u, double* cmax) {
  WB_CUDA(cudaSetDevice(h->dev));
  WB_CUDA(cudaMemcpyAsync(h->u, u, sizeof(double) * NV * h->p.nx, cudaMemcpyHostToDevice, h->stream));
  k1_max_speed<<<1, 256, 0, h->stream>>>(h->u, h->p, h->ctrl, 0);
  WB_LAUNCH_CHECK();
  WB_CUDA(cudaMemcpyAsync(h->h_ctrl, h->ctrl, sizeof(Ctrl1), cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  *cmax = h->h_ctrl->cmax;
  return WB_OK;
}
// evolve (fvm.f90:56-76; benchmark_1d.f90:200-261 incl. the FVM/EQL second stage evaluated at u, :228/:236)
static int h1_evolve(Handle1* h, double* u, const double* w_eq, double tend, int max_iter, int* iters, double* t_out,
                     double* dt_out) {
  WB_CUDA(cudaSetDevice(h->dev));
  const int n = NV * h->p.nx;
  size_t fb = sizeof(double) * n;
  WB_CUDA(cudaMemcpyAsync(h->u, u, fb, cudaMemcpyHostToDevice, h->stream));
  if (w_eq) WB_CUDA(cudaMemcpyAsync(h->weq, w_eq, fb, cudaMemcpyHostToDevice, h->stream));
  k1_ctrl_init<<<1, 1, 0, h->stream>>>(h->ctrl, tend, max_iter, 1);
  WB_LAUNCH_CHECK();
  const int scheme = h->p.solver;
  const bool second_stage_at_u = !h->is_fvm && (scheme == 1 || scheme == 2);
  dim3 b(128), g((n + 127) / 128);
  int it = 0;
  double t = 0.0, dt = 0.0;
  for (;;) {
    if (!(t < tend) || (max_iter >= 0 && it >= max_iter)) break;
    for (int s = 0; s < 32; ++s) {
      k1_max_speed<<<1, 256, 0, h->stream>>>(h->u, h->p, h->ctrl, 1);
      WB_LAUNCH_CHECK();
      WB_CHECK(h1_rhs(h, h->u, scheme, true));
      k1_axpy<<<g, b, 0, h->stream>>>(h->w1, h->u, nullptr, h->dudt, n, 1, h->ctrl);
      WB_LAUNCH_CHECK();
      WB_CHECK(h1_rhs(h, second_stage_at_u ? h->u : h->w1, scheme, true));
      k1_axpy<<<g, b, 0, h->stream>>>(h->u, h->u, h->w1, h->dudt, n, 2, h->ctrl);
      WB_LAUNCH_CHECK();
      k1_advance<<<1, 1, 0, h->stream>>>(h->ctrl);
      WB_LAUNCH_CHECK();
    }
    WB_CUDA(cudaMemcpyAsync(h->h_ctrl, h->ctrl, sizeof(Ctrl1), cudaMemcpyDeviceToHost, h->stream));
    WB_CUDA(cudaStreamSynchronize(h->stream));
    it = h->h_ctrl->iter; t = h->h_ctrl->t; dt = h->h_ctrl->dt;
  }
  WB_CUDA(cudaMemcpyAsync(u, h->u, fb, cudaMemcpyDeviceToHost, h->stream));
  WB_CUDA(cudaStreamSynchronize(h->stream));
  if (iters) *iters = it;
  if (t_out) *t_out = t;
  if (dt_out) *dt_out = dt;
  return WB_OK;
}

}}  // namespace wb::fv1d

using namespace wb;
using namespace wb::fv1d;

struct wb_fvm1d { Handle1* h; };
struct wb_fv1d { Handle1* h; };

extern "C" {

// ---------------------------------------------------------------- fvm.f90
int wb_fvm1d_create(wb_fvm1d** out, const wb_fvm1d_params* p) {
  if (!out || !p) { set_error("null argument"); return WB_ERR_ARG; }
  *out = nullptr;
  WB_REQUIRE(p->nvar == 3, "nvar must be 3 (got %d)", p->nvar);
  WB_REQUIRE(p->nx >= 3, "nx must be >= 3");
  WB_REQUIRE(p->bc == 1 || p->bc == 2, "bc must be 1 (periodic) or 2 (zero gradient): fvm.f90 indexes out of bounds otherwise");
  WB_REQUIRE(p->source == 1 || p->source == 2, "source must be 1 or 2");
  WB_REQUIRE(p->gamma > 1.0 && p->boxlen > 0 && p->n >= 0, "gamma>1, boxlen>0, n>=0 required");
  P1 q{};
  q.nx = p->nx; q.bc = p->bc; q.source = p->source; q.nequilibrium = 0; q.solver = 0;
  q.gamma = p->gamma; q.boxlen = p->boxlen; q.dt_den = (2.0 * (double)p->n + 1.0);
  Handle1* h = nullptr;
  WB_CHECK(h1_create(&h, q, p->device, true));
  *out = new wb_fvm1d{h};
  return WB_OK;
}
int wb_fvm1d_destroy(wb_fvm1d* s) { if (s) { h1_destroy(s->h); delete s; } return WB_OK; }
int wb_fvm1d_compute_update(wb_fvm1d* s, const double* u, double* dudt) {
  if (!s || !u || !dudt) { set_error("null argument"); return WB_ERR_ARG; }
  return h1_update_host(s->h, u, nullptr, dudt, 0);
}
int wb_fvm1d_compute_max_speed(wb_fvm1d* s, const double* u, double* cmax) {
  if (!s || !u || !cmax) { set_error("null argument"); return WB_ERR_ARG; }
  return h1_max_speed_host(s->h, u, cmax);
}
int wb_fvm1d_evolve(wb_fvm1d* s, double* u, double tend, int max_iter, int* iters, double* t_out, double* dt_out) {
  if (!s || !u) { set_error("null argument"); return WB_ERR_ARG; }
  return h1_evolve(s->h, u, nullptr, tend, max_iter, iters, t_out, dt_out);
}

// ---------------------------------------------------------------- benchmark_1d.f90
int wb_fv1d_create(wb_fv1d** out, const wb_fv1d_params* p) {
  if (!out || !p) { set_error("null argument"); return WB_ERR_ARG; }
  *out = nullptr;
  WB_REQUIRE(p->nvar == 3, "nvar must be 3 (got %d)", p->nvar);
  WB_REQUIRE(p->nx >= 4, "nx must be >= 4");
  WB_REQUIRE(p->bc >= 1 && p->bc <= 3, "bc must be 1..3");
  WB_REQUIRE(p->nequilibrium >= 1 && p->nequilibrium <= 3, "nequilibrium must be 1..3");
  WB_REQUIRE(p->solver >= 1 && p->solver <= 3, "solver must be 1 ('FVM'), 2 ('EQL') or 3 ('WB1')");
  WB_REQUIRE(p->gamma > 1.0 && p->boxlen > 0, "gamma>1, boxlen>0 required");
  P1 q{};
  q.nx = p->nx; q.bc = p->bc; q.source = 0; q.nequilibrium = p->nequilibrium; q.solver = p->solver;
  q.gamma = p->gamma; q.boxlen = p->boxlen; q.dt_den = (2.0 * (double)1 + 1.0);
  Handle1* h = nullptr;
  WB_CHECK(h1_create(&h, q, p->device, false));
  *out = new wb_fv1d{h};
  return WB_OK;
}
int wb_fv1d_destroy(wb_fv1d* s) { if (s) { h1_destroy(s->h); delete s; } return WB_OK; }
int wb_fv1d_compute_update(wb_fv1d* s, const double* u, const double* w_eq, double* dudt) {
  if (!s || !u || !w_eq || !dudt) { set_error("null argument"); return WB_ERR_ARG; }
  return h1_update_host(s->h, u, w_eq, dudt, 2);
}
int wb_fv1d_compute_update_fvm(wb_fv1d* s, const double* u, const double* w_eq, double* dudt) {
  if (!s || !u || !w_eq || !dudt) { set_error("null argument"); return WB_ERR_ARG; }
  return h1_update_host(s->h, u, w_eq, dudt, 1);
}
int wb_fv1d_compute_update_sr(wb_fv1d* s, const double* u, const double* w_eq, double* dudt) {
  if (!s || !u || !dudt) { set_error("null argument"); return WB_ERR_ARG; }
  return h1_update_host(s->h, u, w_eq, dudt, 3);
}
int wb_fv1d_compute_max_speed(wb_fv1d* s, const double* u, double* cmax) {
  if (!s || !u || !cmax) { set_error("null argument"); return WB_ERR_ARG; }
  return h1_max_speed_host(s->h, u, cmax);
}
int wb_fv1d_evolve(wb_fv1d* s, double* u, const double* w_eq, double tend, int max_iter, int* iters, double* t_out,
                   double* dt_out) {
  if (!s || !u || !w_eq) { set_error("null argument"); return WB_ERR_ARG; }
  return h1_evolve(s->h, u, w_eq, tend, max_iter, iters, t_out, dt_out);
}

}  // extern "C"
