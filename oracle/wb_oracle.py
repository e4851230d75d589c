"""ORACLE -- TEST INFRASTRUCTURE ONLY (ctypes binding of oracle/libwb_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product (fvm-source-wb_b200/) never does.

Array convention: the reference's Fortran u(nvar,nx,ny) is passed as a C-contiguous numpy array of
shape (ny, nx, nvar) -- same bytes.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libwb_oracle.so")


def build(force=False):
    """Compile the C restatement (gcc, -ffp-contract=off)."""
    if force or not os.path.exists(_LIB) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
        for f in os.listdir(_HERE) if f.endswith((".c", ".h", "Makefile"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None
_dp = C.POINTER(C.c_double)


def _ptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class FV2DParams(C.Structure):
    """parameters_2d.f90:3-21."""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nequilibrium", C.c_int), ("gamma", C.c_double),
                ("boxlen_x", C.c_double), ("boxlen_y", C.c_double), ("cfl", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_get_max_threads.restype = C.c_int
    return _lib


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(n))


def max_threads():
    return lib().orc_get_max_threads()


F32 = lambda v: float(np.float32(v))  # noqa: E731  real(4) literal promoted to real(8)


def fv2d_params(nx, ny, nequilibrium=2, gamma=F32(1.4), boxlen_x=1.0, boxlen_y=1.0, cfl=0.5):
    return FV2DParams(nx, ny, nequilibrium, gamma, boxlen_x, boxlen_y, cfl)


# ---------------------------------------------------------------- 2D FV (benchmark_2d.f90)
def fv2d_get_coords(p):
    x = np.empty((p.ny, p.nx)); y = np.empty((p.ny, p.nx))
    lib().orc_fv2d_get_coords(C.c_int(p.nx), C.c_int(p.ny), C.c_double(p.boxlen_x),
                              C.c_double(p.boxlen_y), _ptr(x), _ptr(y))
    return x, y


def fv2d_get_equilibrium_solution(p, x, y):
    w = np.empty(x.shape + (4,))
    lib().orc_fv2d_get_equilibrium_solution(C.c_int(p.nequilibrium), _ptr(x), _ptr(y), _ptr(w),
                                            C.c_long(x.size))
    return w


def fv2d_get_initial_conditions(p, ninit, x, y, eta=F32(0.00001)):
    u = np.empty(x.shape + (4,))
    lib().orc_fv2d_get_initial_conditions(C.c_int(ninit), C.c_double(eta), C.c_double(p.gamma),
                                          _ptr(x), _ptr(y), _ptr(u), C.c_long(x.size))
    return u


def fv2d_compute_primitive(p, u):
    w = np.empty_like(u)
    lib().orc_fv2d_compute_primitive(_ptr(u), _ptr(w), C.c_double(p.gamma), C.c_long(u.size // 4))
    return w


def fv2d_compute_conservative(p, w):
    u = np.empty_like(w)
    lib().orc_fv2d_compute_conservative(_ptr(w), _ptr(u), C.c_double(p.gamma), C.c_long(w.size // 4))
    return u


def fv2d_compute_max_speed(p, u):
    c = C.c_double(0)
    lib().orc_fv2d_compute_max_speed(C.byref(p), _ptr(u), C.byref(c))
    return c.value


def fv2d_compute_update_exact(p, u, w_eq):
    dudt = np.empty_like(u)
    lib().orc_fv2d_compute_update_exact(C.byref(p), _ptr(u), _ptr(w_eq), _ptr(dudt))
    return dudt


def fv2d_compute_update(p, u, w_eq):
    dudt = np.empty_like(u)
    lib().orc_fv2d_compute_update(C.byref(p), _ptr(u), _ptr(w_eq), _ptr(dudt))
    return dudt


def fv2d_evolve(p, u, w_eq, tend, max_iter=-1):
    """Returns (u_new, iters, t, last_dt, last_cmax); u is not modified."""
    u = np.array(u, copy=True)
    it = C.c_int(0); t = C.c_double(0); dt = C.c_double(0); cm = C.c_double(0)
    lib().orc_fv2d_evolve(C.byref(p), _ptr(u), _ptr(w_eq), C.c_double(tend), C.c_int(max_iter),
                          C.byref(it), C.byref(t), C.byref(dt), C.byref(cm))
    return u, it.value, t.value, dt.value, cm.value


# ---------------------------------------------------------------- 2D DG (2d/benchmark_2d_dg.f90, 2d/legendre.f90, 2d/limiters.f90)
class DG2DParams(C.Structure):
    """2d/parameters_dg_2d.f90:3-35."""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("mx", C.c_int), ("my", C.c_int), ("bc", C.c_int),
                ("source", C.c_int), ("grad_phi_case", C.c_int), ("flux_id", C.c_int), ("limiter_id", C.c_int),
                ("solver_id", C.c_int), ("ninit", C.c_int), ("gamma", C.c_double), ("boxlen_x", C.c_double),
                ("boxlen_y", C.c_double), ("cfl", C.c_double), ("eps", C.c_double), ("M", C.c_double),
                ("eta", C.c_double)]


LIMITERS = {"none": 0, "ONP": 1, "HIO": 2, "1OR": 3, "LOW": 4, "POS": 5, "PO3": 6}
SOLVERS = {"RK4": 1, "SS4": 2, "EQL": 3, "DEB": 4}
FLUXES = {"llf": 0, "llf1": 1, "hll2": 2, "hllc": 3}   # 'llf' is the shipped default that matches no branch (numerical flux stays 0)


def dg2d_params(nx=8, ny=8, mx=2, my=2, bc=1, source=1, grad_phi_case=2, flux="llf1", limiter="ONP", solver="RK4",
                ninit=1, gamma=F32(1.4), boxlen_x=1.0, boxlen_y=1.0, cfl=F32(0.2), eps=F32(1e-10), M=0.0, eta=F32(0.1)):
    return DG2DParams(nx, ny, mx, my, bc, source, grad_phi_case, FLUXES[flux], LIMITERS[limiter], SOLVERS[solver], ninit,
                      gamma, boxlen_x, boxlen_y, cfl, eps, M, eta)


def _dg_shape(p):
    return (p.my, p.mx, p.ny, p.nx)


def dg2d_basis(p):
    gll = (2 * (p.mx - 1) + 3) // 2
    xq = np.zeros(p.mx); wx = np.zeros(p.mx); xg = np.zeros(max(gll, 1)); wg = np.zeros(max(gll, 1))
    lib().orc_dg2d_basis_tables(C.byref(p), _ptr(xq), _ptr(wx), _ptr(xg), _ptr(wg))
    return xq, wx, xg, wg


def dg2d_legendre(x, n):
    f = lib().orc_dg2d_legendre; f.restype = C.c_double
    v = C.c_double(x)
    return f(C.byref(v), C.c_int(n))


def dg2d_legendre_prime(x, n):
    f = lib().orc_dg2d_legendre_prime; f.restype = C.c_double
    v = C.c_double(x)
    return f(C.byref(v), C.c_int(n))


def dg2d_get_coords(p):
    x = np.empty(_dg_shape(p)); y = np.empty(_dg_shape(p))
    lib().orc_dg2d_get_coords(C.byref(p), _ptr(x), _ptr(y))
    return x, y


def dg2d_get_initial_conditions(p, x, y):
    u = np.empty(_dg_shape(p) + (4,))
    lib().orc_dg2d_get_initial_conditions(C.byref(p), _ptr(x), _ptr(y), _ptr(u))
    return u


def dg2d_compute_primitive(p, u):
    """compute_primitive  2d/benchmark_2d_dg.f90:891-902 on an array of states (..., 4)"""
    u = np.ascontiguousarray(u)
    w = np.empty_like(u)
    lib().orc_dg2d_compute_primitive(C.byref(p), _ptr(u), _ptr(w), C.c_long(u.size // 4))
    return w


def dg2d_get_modes_from_nodes(p, nodes):
    u = np.empty_like(nodes)
    lib().orc_dg2d_get_modes_from_nodes(C.byref(p), _ptr(nodes), _ptr(u))
    return u


def dg2d_get_nodes_from_modes(p, modes):
    u = np.empty_like(modes)
    lib().orc_dg2d_get_nodes_from_modes(C.byref(p), _ptr(modes), _ptr(u))
    return u


def dg2d_num_flux(p, ul, ur, flag):
    """compute_num_flux at one face point (flag 1 = x face, 2 = y face)"""
    a = np.ascontiguousarray(ul, dtype=np.float64); b = np.ascontiguousarray(ur, dtype=np.float64)
    nf = np.zeros(4)
    lib().orc_dg2d_num_flux(C.byref(p), _ptr(a), _ptr(b), C.c_int(flag), _ptr(nf))
    return nf


def dg2d_compute_update(p, modes, x, y):
    d = np.empty_like(modes)
    lib().orc_dg2d_compute_update(C.byref(p), _ptr(modes), _ptr(x), _ptr(y), _ptr(d))
    return d


def dg2d_compute_max_speed(p, modes):
    a = [C.c_double() for _ in range(4)]
    lib().orc_dg2d_compute_max_speed(C.byref(p), _ptr(modes), *[C.byref(v) for v in a])
    return tuple(v.value for v in a)   # cs_max, v_xmax, v_ymax, speed_max


def dg2d_apply_limiter(p, modes):
    u = np.array(modes, copy=True)
    lib().orc_dg2d_apply_limiter(C.byref(p), _ptr(u))
    return u


def dg2d_compute_error(p, u_nodes, u_init):
    """compute_error :23-89 -> (lmax[4], l1[4], l2[4]); l2 is the accumulator before the sqrt"""
    a = np.zeros(4); b = np.zeros(4); c = np.zeros(4)
    lib().orc_dg2d_compute_error(C.byref(p), _ptr(u_nodes), _ptr(u_init), _ptr(a), _ptr(b), _ptr(c))
    return a, b, c


def dg2d_evolve_modes(p, modes, x, y, tend, max_iter=-1):
    u = np.array(modes, copy=True)
    it = C.c_int(); t = C.c_double(); dt = C.c_double()
    lib().orc_dg2d_evolve_modes(C.byref(p), _ptr(u), _ptr(x), _ptr(y), C.c_double(tend), C.c_int(max_iter),
                                C.byref(it), C.byref(t), C.byref(dt))
    return u, it.value, t.value, dt.value


def dg2d_evolve(p, nodes, x, y, tend, max_iter=-1):
    u = np.array(nodes, copy=True)
    it = C.c_int(); t = C.c_double(); dt = C.c_double()
    lib().orc_dg2d_evolve(C.byref(p), _ptr(u), _ptr(x), _ptr(y), C.c_double(tend), C.c_int(max_iter),
                          C.byref(it), C.byref(t), C.byref(dt))
    return u, it.value, t.value, dt.value


# ---------------------------------------------------------------- 1D FV (fvm.f90, benchmark_1d.f90)
class FVM1DParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("bc", C.c_int), ("source", C.c_int), ("n", C.c_int), ("gamma", C.c_double),
                ("boxlen", C.c_double)]


def fvm1d_params(nx=200, bc=2, source=2, n=3, gamma=F32(1.4), boxlen=1.0):
    return FVM1DParams(nx, bc, source, n, gamma, boxlen)


def fvm1d_initial_conditions(p, ninit=4):
    u = np.empty((p.nx, 3))
    lib().orc_fvm1d_initial_conditions(C.byref(p), C.c_int(ninit), _ptr(u))
    return u


def fvm1d_compute_update(p, u):
    d = np.empty_like(u)
    lib().orc_fvm1d_compute_update(C.byref(p), _ptr(u), _ptr(d))
    return d


def fvm1d_compute_max_speed(p, u):
    c = C.c_double()
    lib().orc_fvm1d_compute_max_speed(C.byref(p), _ptr(u), C.byref(c))
    return c.value


def fvm1d_evolve(p, u, tend, max_iter=-1):
    u = np.array(u, copy=True)
    it = C.c_int(); t = C.c_double(); dt = C.c_double()
    lib().orc_fvm1d_evolve(C.byref(p), _ptr(u), C.c_double(tend), C.c_int(max_iter), C.byref(it), C.byref(t), C.byref(dt))
    return u, it.value, t.value, dt.value


class FV1DParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("bc", C.c_int), ("nequilibrium", C.c_int), ("solver", C.c_int), ("gamma", C.c_double),
                ("boxlen", C.c_double)]


SOLVERS_1D = {"FVM": 1, "EQL": 2, "WB1": 3}


def fv1d_params(nx=128, bc=2, nequilibrium=2, solver="WB1", gamma=F32(1.4), boxlen=1.0):
    return FV1DParams(nx, bc, nequilibrium, SOLVERS_1D[solver], gamma, boxlen)


def fv1d_get_x(p):
    x = np.empty(p.nx)
    lib().orc_fv1d_get_x(C.byref(p), _ptr(x))
    return x


def fv1d_get_equilibrium_solution(p, x):
    w = np.empty((x.size, 3))
    lib().orc_fv1d_get_equilibrium_solution(C.byref(p), _ptr(x), _ptr(w), C.c_int(x.size))
    return w


def fv1d_get_initial_conditions(p, ninit, x, eta=F32(1e-8)):
    u = np.empty((p.nx, 3))
    lib().orc_fv1d_get_initial_conditions(C.byref(p), C.c_int(ninit), C.c_double(eta), _ptr(x), _ptr(u))
    return u


def _fv1d_upd(fn, p, u, w_eq):
    d = np.empty_like(u)
    getattr(lib(), fn)(C.byref(p), _ptr(u), _ptr(w_eq), _ptr(d))
    return d


def fv1d_compute_update(p, u, w_eq):
    return _fv1d_upd("orc_fv1d_compute_update", p, u, w_eq)


def fv1d_compute_update_fvm(p, u, w_eq):
    return _fv1d_upd("orc_fv1d_compute_update_fvm", p, u, w_eq)


def fv1d_compute_update_sr(p, u, w_eq):
    return _fv1d_upd("orc_fv1d_compute_update_sr", p, u, w_eq)


def fv1d_compute_max_speed(p, u):
    c = C.c_double()
    lib().orc_fv1d_compute_max_speed(C.byref(p), _ptr(u), C.byref(c))
    return c.value


def fv1d_evolve(p, u, w_eq, tend, max_iter=-1):
    u = np.array(u, copy=True)
    it = C.c_int(); t = C.c_double(); dt = C.c_double()
    lib().orc_fv1d_evolve(C.byref(p), _ptr(u), _ptr(w_eq), C.c_double(tend), C.c_int(max_iter), C.byref(it), C.byref(t),
                          C.byref(dt))
    return u, it.value, t.value, dt.value


# ---------------------------------------------------------------- 1D DG (dg_with_source.f90, root legendre.f90)
class DG1DParams(C.Structure):
    _fields_ = [("n", C.c_int), ("nx", C.c_int), ("riemann", C.c_int), ("source", C.c_int), ("ninit", C.c_int),
                ("gamma", C.c_double), ("boxlen", C.c_double), ("pert", C.c_double), ("bc", C.c_int), ("use_limiter", C.c_int)]


def dg1d_params(n=3, nx=128, riemann=2, source=2, ninit=8, gamma=F32(1.4), boxlen=1.0, pert=F32(1e-8), bc=5, use_limiter=0):
    return DG1DParams(n, nx, riemann, source, ninit, gamma, boxlen, pert, bc, use_limiter)


def dg1d_quadrature(p):
    x = np.zeros(p.n); w = np.zeros(p.n)
    lib().orc_dg1d_quadrature(C.byref(p), _ptr(x), _ptr(w))
    return x, w


def dg1d_legendre(x, n):
    f = lib().orc_dg1d_legendre; f.restype = C.c_double
    v = C.c_double(x)
    return f(C.byref(v), C.c_int(n))


def dg1d_setup(p):
    """Returns (uinit nodal IC, u_eq nodal equilibrium, delta_u projected perturbation)  program dg :33-171."""
    shp = (p.nx, p.n, 3)
    a = np.empty(shp); b = np.empty(shp); c = np.empty(shp)
    lib().orc_dg1d_setup(C.byref(p), _ptr(a), _ptr(b), _ptr(c))
    return a, b, c


def dg1d_compute_update_exact_delta(p, delta_u, u_eq):
    d = np.empty_like(delta_u)
    lib().orc_dg1d_compute_update_exact_delta(C.byref(p), _ptr(delta_u), _ptr(u_eq), _ptr(d))
    return d


def dg1d_compute_max_speed(p, u_nodes):
    c = C.c_double()
    lib().orc_dg1d_compute_max_speed(C.byref(p), _ptr(u_nodes), C.byref(c))
    return c.value


def dg1d_evolve_rki(p, delta_u, u_eq, uinit, tend, max_iter=-1):
    d = np.array(delta_u, copy=True); ui = np.array(uinit, copy=True)
    it = C.c_int(); t = C.c_double(); dt = C.c_double()
    lib().orc_dg1d_evolve_rki(C.byref(p), _ptr(d), _ptr(u_eq), _ptr(ui), C.c_double(tend), C.c_int(max_iter), C.byref(it),
                              C.byref(t), C.byref(dt))
    return d, ui, it.value, t.value, dt.value


def dg1d_project(p, u_nodes):
    m = np.empty_like(u_nodes)
    lib().orc_dg1d_project(C.byref(p), _ptr(u_nodes), _ptr(m))
    return m


def dg1d_compute_update(p, u):
    d = np.empty_like(u)
    lib().orc_dg1d_compute_update(C.byref(p), _ptr(u), _ptr(d))
    return d


def dg1d_limiter(p, u):
    v = np.array(u, copy=True)
    lib().orc_dg1d_limiter(C.byref(p), _ptr(v))
    return v


INTEGRATORS_1D = {"RK1": 1, "RK2": 2, "RK3": 3, "RK4": 4}


def dg1d_evolve_rk(p, integrator, u, delta_u, u_eq, uinit, tend, max_iter=-1):
    uu = np.array(u, copy=True); ui = np.array(uinit, copy=True)
    it = C.c_int(); t = C.c_double(); dt = C.c_double()
    lib().orc_dg1d_evolve_rk(C.byref(p), C.c_int(INTEGRATORS_1D[integrator]), _ptr(uu), _ptr(delta_u), _ptr(u_eq), _ptr(ui),
                             C.c_double(tend), C.c_int(max_iter), C.byref(it), C.byref(t), C.byref(dt))
    return uu, ui, it.value, t.value, dt.value


def dg1d_compute_update_exact(p, u, u_eq_modes):
    d = np.empty_like(u)
    lib().orc_dg1d_compute_update_exact(C.byref(p), _ptr(u), _ptr(u_eq_modes), _ptr(d))
    return d


def dg1d_limiter_tdv(p, u):
    v = np.array(u, copy=True)
    lib().orc_dg1d_limiter_tdv(C.byref(p), _ptr(v))
    return v


def dg1d_limiter_cons(p, u):
    v = np.array(u, copy=True)
    lib().orc_dg1d_limiter_cons(C.byref(p), _ptr(v))
    return v


def dg1d_evolve_w(p, integrator, u, delta_u, u_eq_nodes, u_eq_modes, uinit, tend, max_iter=-1):
    uu = np.array(u, copy=True); dd = np.array(delta_u, copy=True); ui = np.array(uinit, copy=True)
    it = C.c_int(); t = C.c_double(); dt = C.c_double()
    lib().orc_dg1d_evolve_w(C.byref(p), C.c_int({"RKw": 5, "RKe": 6}[integrator]), _ptr(uu), _ptr(dd), _ptr(u_eq_nodes),
                            _ptr(u_eq_modes), _ptr(ui), C.c_double(tend), C.c_int(max_iter), C.byref(it), C.byref(t), C.byref(dt))
    return uu, dd, ui, it.value, t.value, dt.value
