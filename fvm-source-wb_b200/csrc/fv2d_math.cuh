// Pointwise physics of the 2D well-balanced FV path (benchmark_2d.f90), two arithmetic flavours:
//   ref::  reference operation order, IEEE div/sqrt, no FMA  (file is compiled with -fmad=false, so
//          plain C expressions are evaluated exactly as written)
//   fast:: fused arithmetic with Newton reciprocal / rsqrt; same formulas, <= few ulp per operation.
// Both keep the property that makes the scheme well balanced bit for bit: the numerical-flux side
// and the equilibrium-flux side run the SAME instruction sequence on the same inputs.
#pragma once
#include <cuda_runtime.h>

namespace wb { namespace fv2d {

struct Phys {
  double gamma, gm1;       // gamma, gamma-1 (gamma - 1.0 with the reference's promoted literal)
  double dx, dy, odx, ody;
  double hodx, hody;       // 0.5/dx, 0.5/dy (fused kernels carry doubled face fluxes)
  double gm1x2;            // 2*(gamma-1)
  double cfl;
  double rho0, p0, a;      // equilibrium constants: rho = rho0*exp(-a(x+y)), p = p0*exp(-a(x+y))
  double pe1;              // p0/(gamma-1)
  int neq;
};

// ------------------------------------------------------------------ coordinates (benchmark_2d.f90:505-522)
// i, j are 0-based GLOBAL indices; the reference is 1-based: (i-1)*dx -> i*dx, (i-0.5) in real(4).
__device__ __forceinline__ double x_face(int i, double dx) { return (double)i * dx; }
__device__ __forceinline__ double x_cent(int i, double dx) { return (double)((float)(i + 1) - 0.5f) * dx; }

namespace ref {

// benchmark_2d.f90:174-218 (primitives rho, p; velocities are zero in every case)
__device__ __forceinline__ void eq_prim(const Phys& P, double x, double y, double& rho, double& p) {
  if (P.neq == 4) { rho = 0.0; p = 0.0; return; }
  double e = exp(-(P.a) * (x + y));      // case 1: a = 1, -(x+y) == (-1)*(x+y) bit for bit
  rho = P.rho0 * e;
  p = P.p0 * e;
}
// benchmark_2d.f90:159-171
__device__ __forceinline__ void cons(const Phys& P, const double w[4], double u[4]) {
  u[0] = w[0];
  u[1] = w[0] * w[1];
  u[2] = w[0] * w[2];
  u[3] = w[3] / P.gm1 + 0.5 * (w[0] * (w[1] * w[1] + w[2] * w[2]));
}
// benchmark_2d.f90:145-157
__device__ __forceinline__ void prim(const Phys& P, const double u[4], double w[4]) {
  w[0] = u[0];
  w[1] = u[1] / w[0];
  w[2] = u[2] / w[0];
  w[3] = P.gm1 * (u[3] - 0.5 * w[0] * (w[1] * w[1] + w[2] * w[2]));
}
// benchmark_2d.f90:283-295
__device__ __forceinline__ double speed(const Phys& P, const double u[4]) {
  double w[4];
  prim(P, u, w);
  double cs = sqrt(P.gamma * fmax(w[3], 1e-10) / fmax(w[0], 1e-10));
  return sqrt(w[1] * w[1] + w[2] * w[2]) + cs;
}
// benchmark_2d.f90:299-325, one direction
template <int DIR>
__device__ __forceinline__ void flux(const Phys& P, const double u[4], double f[4]) {
  double w[4];
  prim(P, u, w);
  if (DIR == 0) {
    f[0] = w[1] * u[0];
    f[1] = w[1] * u[1] + w[3];
    f[2] = w[0] * w[1] * w[2];
    f[3] = w[1] * u[3] + w[1] * w[3];
  } else {
    f[0] = u[0] * w[2];
    f[1] = u[1] * w[2];
    f[2] = u[2] * w[2] + w[3];
    f[3] = w[2] * u[3] + w[2] * w[3];
  }
}
// benchmark_2d.f90:353-367
template <int DIR>
__device__ __forceinline__ void llf(const Phys& P, const double ul[4], const double ur[4], double fg[4]) {
  double fl[4], fr[4];
  flux<DIR>(P, ul, fl);
  flux<DIR>(P, ur, fr);
  double cl = speed(P, ul), cr = speed(P, ur);
  double cmax = fmax(cl, cr);
#pragma unroll
  for (int v = 0; v < 4; ++v) fg[v] = 0.5 * (fr[v] + fl[v]) + 0.5 * cmax * (ul[v] - ur[v]);
}
// benchmark_2d.f90:327-350 with phi_x = phi_y = 1.
__device__ __forceinline__ void source(const double w[4], double s[4]) {
  s[0] = 0.0;
  s[1] = -w[0] * 1.0;
  s[2] = -w[0] * 1.0;
  s[3] = -w[0] * (w[1] * 1.0 + w[2] * 1.0);
}

}  // namespace ref

namespace fast {

// max of two wave speeds (finite, non-negative): one compare and a select.  fmax() costs nine instructions on sm_100a (there
// is no FP64 min/max instruction: DSETP.MAX plus the NaN and signed-zero fix-ups), this costs three and gives the same
// number for every pair of finite values; a NaN in `b` still propagates, and a NaN state poisons the fluxes anyway.
__device__ __forceinline__ double max_speed2(double a, double b) { return a > b ? a : b; }
// a > b ? a : b through PTX: written in C++ with a constant `b` the compiler recognises max.f64 and emits the nine-instruction
// sequence again.  For finite or NaN a and a finite constant b this IS fmax(a, b).
__device__ __forceinline__ double sel_gt(double a, double b) {
  double d;
  asm("{\n.reg .pred p;\nsetp.gt.f64 p, %1, %2;\nselp.f64 %0, %1, %2, p;\n}" : "=d"(d) : "d"(a), "d"(b));
  return d;
}

// 1/x: MUFU.RCP64H seed (~2^-20) + one cubic Newton step; relative error ~2^-53.
__device__ __forceinline__ double rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  double t = fma(e, e, e);
  return fma(y, t, y);
}
// sqrt(a) for a > 0 (callers clamp): MUFU.RSQ64H seed + cubic step on 1/sqrt, then a*y.
__device__ __forceinline__ double sqrt_pos(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double t = a * y;
  double h = fma(-t, y, 1.0);
  double p = fma(0.375, h, 0.5);
  p = p * h;
  double s = fma(t, p, t);        // a*y*(1 + h/2 + 3h^2/8)
  return s;
}

// Directional state evaluation in (normal, tangential) momentum form.
//   f = (mn, vn*mn + p, vn*mt, vn*(E+p)),  spd = |v| + sqrt(gamma*max(p,1e-10)/max(rho,1e-10))
struct Eval { double f0, fn, ft, f3, spd; };
__device__ __forceinline__ Eval eval_state(const Phys& P, double rho, double mn, double mt, double E) {
  Eval o;
  double r = rcp(rho);
  double vn = mn * r, vt = mt * r;
  double q = fma(vt, vt, vn * vn);
  double p = P.gm1 * fma(-0.5 * rho, q, E);
  double pm = sel_gt(p, 1e-10);      // == fmax(p, 1e-10) for every p, NaN included
  double rm = (rho >= 1e-10) ? r : 1e10;
  double c2 = (P.gamma * pm) * rm;
  double cs = sqrt_pos(c2);
  double vm = sqrt_pos(q + 1e-300);        // q == 0 gives 1e-150, absorbed by "+ cs"; q > 1e-284 is unchanged
  o.spd = vm + cs;
  o.f0 = mn;
  o.fn = fma(vn, mn, p);
  o.ft = vn * mt;
  o.f3 = vn * (E + p);
  return o;
}
// The same evaluation for N independent states in lock-step: every sub-step is issued for all N states before the
// next one, so the instruction stream carries N independent dependency chains.  Bit-identical to eval_state.
// EXACT = false leaves the floors max(p,1d-10), max(rho,1d-10) out and clears `ok` when any state is close enough to a
// floor for them to matter (tested on the integer pipe: sign and exponent live in the high word, hi >= hi(1e-10)+1
// implies value > 1e-10); the caller then re-evaluates with EXACT = true.  Physical states never get there, and the
// kernel is bound by issue slots (an FP64 instruction takes two), so the floors cost ~5 % if evaluated always.
template <int N, bool EXACT>
__device__ __forceinline__ void eval_states(const Phys& P, const double (&rho)[N], const double (&mn)[N], const double (&mt)[N],
                                            const double (&E)[N], Eval (&o)[N], bool& ok) {
  double r[N], y[N], e[N], t[N], vn[N], vt[N], q[N], p[N], c2[N], cs[N], vm[N], h[N], w[N], qa[N];
#pragma unroll
  for (int k = 0; k < N; ++k) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[k]) : "d"(rho[k]));
#pragma unroll
  for (int k = 0; k < N; ++k) e[k] = fma(-rho[k], y[k], 1.0);
#pragma unroll
  for (int k = 0; k < N; ++k) t[k] = fma(e[k], e[k], e[k]);
#pragma unroll
  for (int k = 0; k < N; ++k) r[k] = fma(y[k], t[k], y[k]);
#pragma unroll
  for (int k = 0; k < N; ++k) { vn[k] = mn[k] * r[k]; vt[k] = mt[k] * r[k]; }
#pragma unroll
  for (int k = 0; k < N; ++k) q[k] = fma(vt[k], vt[k], vn[k] * vn[k]);
#pragma unroll
  for (int k = 0; k < N; ++k) p[k] = P.gm1 * fma(-0.5 * rho[k], q[k], E[k]);
  if (EXACT) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double pm = sel_gt(p[k], 1e-10);
      const double rm = (rho[k] >= 1e-10) ? r[k] : 1e10;
      c2[k] = (P.gamma * pm) * rm;
    }
  } else {
    constexpr int HI_OK = 0x3DDB7CDF + 1;      // high word of 1e-10 is 0x3DDB7CDF
    int lo = 0x7fffffff;                       // signed minimum of the high words: negative values are caught as well
#pragma unroll
    for (int k = 0; k < N; ++k) {
      lo = min(lo, min(__double2hiint(p[k]), __double2hiint(rho[k])));
      c2[k] = (P.gamma * p[k]) * r[k];
    }
    ok = ok && (lo >= HI_OK);
  }
#pragma unroll
  for (int k = 0; k < N; ++k) qa[k] = q[k] + 1e-300;
  // two square roots per state, all 2N in lock-step (sqrt_pos)
#pragma unroll
  for (int k = 0; k < N; ++k) {
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[k]) : "d"(c2[k]));
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(w[k]) : "d"(qa[k]));
  }
#pragma unroll
  for (int k = 0; k < N; ++k) { t[k] = c2[k] * y[k]; e[k] = qa[k] * w[k]; }
#pragma unroll
  for (int k = 0; k < N; ++k) { h[k] = fma(-t[k], y[k], 1.0); w[k] = fma(-e[k], w[k], 1.0); }
#pragma unroll
  for (int k = 0; k < N; ++k) { y[k] = fma(0.375, h[k], 0.5); vm[k] = fma(0.375, w[k], 0.5); }
#pragma unroll
  for (int k = 0; k < N; ++k) { y[k] = y[k] * h[k]; vm[k] = vm[k] * w[k]; }
#pragma unroll
  for (int k = 0; k < N; ++k) { cs[k] = fma(t[k], y[k], t[k]); vm[k] = fma(e[k], vm[k], e[k]); }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    o[k].spd = vm[k] + cs[k];
    o[k].f0 = mn[k];
    o[k].fn = fma(vn[k], mn[k], p[k]);
    o[k].ft = vn[k] * mt[k];
    o[k].f3 = vn[k] * (E[k] + p[k]);
  }
}

// max wave speed of a state (stage-2 CFL reduction)
__device__ __forceinline__ double speed(const Phys& P, double rho, double mx, double my, double E) {
  double r = rcp(rho);
  double vx = mx * r, vy = my * r;
  double q = fma(vy, vy, vx * vx);
  double p = P.gm1 * fma(-0.5 * rho, q, E);
  double pm = sel_gt(p, 1e-10);      // == fmax(p, 1e-10) for every p, NaN included
  double rm = (rho >= 1e-10) ? r : 1e10;
  double cs = sqrt_pos((P.gamma * pm) * rm);
  double vm = sqrt_pos(q + 1e-300);
  return vm + cs;
}

}  // namespace fast

}}  // namespace wb::fv2d
