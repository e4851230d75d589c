"""wbeuler -- host-side mirror of the reference's update routines over the C-ABI in include/wbeuler.h.

The reference (hanveiga/fvm-source-wb) is serial Fortran whose "interface" is a set of external
subroutines (`compute_update_exact`, `compute_max_speed`, `evolve`, ...).  This package exposes the
same names with the same argument meaning on top of ``libwbeuler.so`` (hand-written FP64 CUDA for
sm_100a).  There is no CPU fallback: if the shared library is missing, or no B200 is visible, every
call raises.

Array convention: the reference's ``u(nvar,nx,ny)`` is a C-contiguous float64 numpy array of shape
``(ny, nx, nvar)`` (identical bytes).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WBEULER_LIB") or os.path.join(_HERE, "libwbeuler.so")   # override: A/B builds in development

_lib = None
_dp = C.POINTER(C.c_double)


class WBError(RuntimeError):
    pass


def lib():
    """Load libwbeuler.so (raises if it has not been built: `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WBError(f"{LIB_PATH} is missing: the CUDA library has not been built "
                          "(run __graft_entry__.build()); there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.wb_last_error.restype = C.c_char_p
        _lib.wb_version.restype = C.c_char_p
        _lib.wb_kernel_launch_count.restype = C.c_longlong
    return _lib


def _check(status):
    if status != 0:
        raise WBError(f"wbeuler error {status}: {lib().wb_last_error().decode()}")


def version():
    return lib().wb_version().decode()


def kernel_launch_count():
    return int(lib().wb_kernel_launch_count())


def nccl_get_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().wb_nccl_get_unique_id(buf))
    return buf.raw


def _ptr(a, shape=None):
    """Pointer to a C-contiguous float64 array; with `shape`, the array must have exactly that shape (the library copies
    prod(shape) doubles from the pointer: a wrong-sized array would be an out-of-bounds host read)."""
    if not isinstance(a, np.ndarray):
        raise WBError(f"expected a numpy array, got {type(a).__name__}")
    if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
        raise WBError("arrays must be C-contiguous float64")
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise WBError(f"array of shape {tuple(a.shape)} where {tuple(shape)} is expected")
    return a.ctypes.data_as(_dp)


def F32(v):
    """A real(4) literal of the reference promoted to real(8) (its Makefiles set no -fdefault-real-8)."""
    return float(np.float32(v))


class FV2DParams(C.Structure):
    """wb_fv2d_params (include/wbeuler.h); defaults are parameters_2d.f90:3-21."""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nvar", C.c_int), ("nequilibrium", C.c_int),
                ("gamma", C.c_double), ("boxlen_x", C.c_double), ("boxlen_y", C.c_double),
                ("cfl", C.c_double), ("arith", C.c_int), ("device", C.c_int), ("rank", C.c_int),
                ("nranks", C.c_int)]


class FV2D:
    """2D well-balanced finite volumes (benchmark_2d.f90).  Methods carry the reference's names."""

    def __init__(self, nx, ny, nequilibrium=2, gamma=F32(1.4), boxlen_x=1.0, boxlen_y=1.0, cfl=0.5,
                 arith=0, device=-1, rank=0, nranks=1):
        self.params = FV2DParams(nx, ny, 4, nequilibrium, gamma, boxlen_x, boxlen_y, cfl, arith, device,
                                 rank, nranks)
        self._h = C.c_void_p()
        _check(lib().wb_fv2d_create(C.byref(self._h), C.byref(self.params)))
        j0 = C.c_int(); nr = C.c_int()
        _check(lib().wb_fv2d_local_rows(self._h, C.byref(j0), C.byref(nr)))
        self.j0, self.nyl = j0.value, nr.value
        self.nx, self.ny = nx, ny

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().wb_fv2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def local_shape(self):
        return (self.nyl, self.nx, 4)

    def set_stream(self, cuda_stream_ptr):
        _check(lib().wb_fv2d_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def comm_init(self, unique_id_bytes):
        _check(lib().wb_fv2d_comm_init(self._h, C.c_char_p(unique_id_bytes)))

    def exchange_kind(self):
        """'p2p' (ghost rows stored into peer memory by the stage kernel), 'nccl' (send/recv) or 'none' (single rank)"""
        f = lib().wb_fv2d_exchange_kind
        f.restype = C.c_char_p
        f.argtypes = [C.c_void_p]
        return f(self._h).decode()

    # -- the reference's routines -----------------------------------------------------------------
    def compute_update_exact(self, u, w_eq):
        """compute_update_exact(u,w_eq,dudt)  benchmark_2d.f90:465-618"""
        dudt = np.empty(self.local_shape)
        _check(lib().wb_fv2d_compute_update_exact(self._h, _ptr(u, self.local_shape), _ptr(w_eq, self.local_shape), _ptr(dudt)))
        return dudt

    def compute_update(self, u, w_eq):
        """compute_update(u,w_eq,dudt)  benchmark_2d.f90:370-463 (plain scheme)"""
        dudt = np.empty(self.local_shape)
        _check(lib().wb_fv2d_compute_update(self._h, _ptr(u, self.local_shape), _ptr(w_eq, self.local_shape), _ptr(dudt)))
        return dudt

    def compute_max_speed(self, u):
        """compute_max_speed(u,cmax)  benchmark_2d.f90:264-279"""
        c = C.c_double()
        _check(lib().wb_fv2d_compute_max_speed(self._h, _ptr(u, self.local_shape), C.byref(c)))
        return c.value

    def evolve(self, u, w_eq, tend, max_iter=-1):
        """evolve(u,u_eq)  benchmark_2d.f90:221-260.  Returns (u_new, iters, t, last_dt)."""
        u = np.array(u, dtype=np.float64, order="C", copy=True)
        it = C.c_int(); t = C.c_double(); dt = C.c_double()
        _check(lib().wb_fv2d_evolve(self._h, _ptr(u, self.local_shape), _ptr(w_eq, self.local_shape), C.c_double(tend), C.c_int(max_iter),
                                    C.byref(it), C.byref(t), C.byref(dt)))
        return u, it.value, t.value, dt.value

    def get_initial_conditions(self, ninit, eta=F32(0.00001)):
        """get_initial_conditions + get_equilibrium_solution at centres (benchmark_2d.f90:45-113,:174-218).
        Returns (u, w_eq) for the local rows."""
        u = np.empty(self.local_shape); w = np.empty(self.local_shape)
        _check(lib().wb_fv2d_get_initial_conditions(self._h, C.c_int(ninit), C.c_double(eta), _ptr(u, self.local_shape), _ptr(w)))
        return u, w

    # -- resident path ------------------------------------------------------------------------------
    def upload(self, u, w_eq):
        _check(lib().wb_fv2d_upload(self._h, _ptr(u, self.local_shape), _ptr(w_eq, self.local_shape)))

    def init_device(self, ninit, eta=F32(0.00001)):
        _check(lib().wb_fv2d_init_device(self._h, C.c_int(ninit), C.c_double(eta)))

    def step_async(self, nsteps, tend=1e300):
        _check(lib().wb_fv2d_step_async(self._h, C.c_int(nsteps), C.c_double(tend)))

    def sync(self):
        """Returns (iters, t, last_dt, last_cmax)."""
        it = C.c_int(); t = C.c_double(); dt = C.c_double(); cm = C.c_double()
        _check(lib().wb_fv2d_sync(self._h, C.byref(it), C.byref(t), C.byref(dt), C.byref(cm)))
        return it.value, t.value, dt.value, cm.value

    def download(self):
        u = np.empty(self.local_shape)
        _check(lib().wb_fv2d_download(self._h, _ptr(u, self.local_shape)))
        return u

    def reset_clock(self):
        _check(lib().wb_fv2d_reset_clock(self._h))

    def output_file(self, path, wait=True):
        """output_file(x,y,u,filen)  benchmark_2d.f90:115-143 of the resident state (asynchronous writer; wait=False returns at once)"""
        _check(lib().wb_fv2d_output_file(self._h, C.c_char_p(os.fsencode(path))))
        if wait:
            self.output_wait()

    def output_wait(self):
        _check(lib().wb_fv2d_output_wait(self._h))


# ================================================================================================== 2D DG
class DG2DParams(C.Structure):
    """wb_dg2d_params (include/wbeuler.h); defaults follow 2d/parameters_dg_2d.f90:3-35."""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("mx", C.c_int), ("my", C.c_int), ("nvar", C.c_int), ("bc", C.c_int),
                ("source", C.c_int), ("grad_phi_case", C.c_int), ("flux_id", C.c_int), ("limiter_id", C.c_int),
                ("solver_id", C.c_int), ("ninit", C.c_int), ("gamma", C.c_double), ("boxlen_x", C.c_double),
                ("boxlen_y", C.c_double), ("cfl", C.c_double), ("eps", C.c_double), ("M", C.c_double), ("device", C.c_int),
                ("arith", C.c_int), ("rank", C.c_int), ("nranks", C.c_int)]


LIMITERS = {"none": 0, "ONP": 1, "HIO": 2, "1OR": 3, "LOW": 4, "POS": 5, "PO3": 6}     # limiter_type (2d/benchmark_2d_dg.f90:1516-1555)
SOLVERS = {"RK4": 1, "SS4": 2, "EQL": 3, "DEB": 4}                 # solver (:672-747)
FLUXES = {"llf": 0, "llf1": 1, "hll2": 2, "hllc": 3}                                     # flux_type; 'llf' is the shipped value that matches no branch


class DG2D:
    """2D modal DG (2d/benchmark_2d_dg.f90).  Arrays: u(nvar,nx,ny,mx,my) == numpy (my, mx, ny, nx, 4)."""

    def __init__(self, nx=8, ny=8, mx=2, my=2, bc=1, source=1, grad_phi_case=2, flux="llf1", limiter="ONP", solver="RK4",
                 ninit=1, gamma=F32(1.4), boxlen_x=1.0, boxlen_y=1.0, cfl=F32(0.2), eps=F32(1e-10), M=0.0, device=-1, arith=0,
                 rank=0, nranks=1):
        self.params = DG2DParams(nx, ny, mx, my, 4, bc, source, grad_phi_case, FLUXES[flux], LIMITERS[limiter],
                                 SOLVERS[solver], ninit, gamma, boxlen_x, boxlen_y, cfl, eps, M, device, arith, rank, nranks)
        self._h = C.c_void_p()
        _check(lib().wb_dg2d_create(C.byref(self._h), C.byref(self.params)))
        j0 = C.c_int(); nr = C.c_int()
        _check(lib().wb_dg2d_local_rows(self._h, C.byref(j0), C.byref(nr)))
        self.j0, self.nrows = j0.value, nr.value             # this rank's slab: global rows [j0, j0 + nrows)
        self.shape = (my, mx, self.nrows, nx, 4)             # host arrays hold the rank's own rows

    def comm_init(self, unique_id_bytes):
        """slab mode: wire the NCCL communicator (every rank, same 128-byte id)"""
        _check(lib().wb_dg2d_comm_init(self._h, C.c_char_p(unique_id_bytes)))

    def exchange_kind(self):
        """'p2p' (ghost rows stored into peer memory by the fused stage kernel), 'nccl' (pack, send/recv, unpack) or 'none'"""
        f = lib().wb_dg2d_exchange_kind
        f.restype = C.c_char_p
        f.argtypes = [C.c_void_p]
        return f(self._h).decode()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().wb_dg2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_ptr):
        _check(lib().wb_dg2d_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def stage_kernel(self):
        """name of the RK-stage kernel this handle launches ("split", "tma", "march", "fast", "reference")"""
        f = lib().wb_dg2d_stage_kernel
        f.restype = C.c_char_p
        f.argtypes = [C.c_void_p]
        return f(self._h).decode()

    def quadrature(self):
        """gl_quadrature(x_quad, w_quad, mx)  2d/legendre.f90:77-108"""
        x = np.zeros(self.params.mx); w = np.zeros(self.params.mx)
        _check(lib().wb_dg2d_quadrature(self._h, _ptr(x), _ptr(w)))
        return x, w

    def get_modes_from_nodes(self, nodes):
        """2d/benchmark_2d_dg.f90:497-542"""
        out = np.empty(self.shape)
        _check(lib().wb_dg2d_get_modes_from_nodes(self._h, _ptr(nodes, self.shape), _ptr(out)))
        return out

    def get_nodes_from_modes(self, modes):
        """2d/benchmark_2d_dg.f90:544-592"""
        out = np.empty(self.shape)
        _check(lib().wb_dg2d_get_nodes_from_modes(self._h, _ptr(modes, self.shape), _ptr(out)))
        return out

    def compute_error(self, u_nodes, u_init_nodes):
        """compute_error(u,x,y,t,u_anal)  2d/benchmark_2d_dg.f90:23-89 -> (lmax[4], l1[4], l2[4] before the sqrt)"""
        a = np.zeros(4); b = np.zeros(4); c = np.zeros(4)
        _check(lib().wb_dg2d_compute_error(self._h, _ptr(u_nodes, self.shape), _ptr(u_init_nodes, self.shape), _ptr(a), _ptr(b), _ptr(c)))
        return a, b, c

    def compute_update(self, modes, x=None, y=None):
        """compute_update(delta_u,x,y,u_eq,dudt)  2d/benchmark_2d_dg.f90:1137-1479"""
        out = np.empty(self.shape)
        _check(lib().wb_dg2d_compute_update(self._h, _ptr(modes, self.shape), _ptr(x, self.shape[:-1]) if x is not None else None,
                                            _ptr(y, self.shape[:-1]) if y is not None else None, _ptr(out)))
        return out

    def apply_limiter(self, modes):
        """apply_limiter(u)  2d/benchmark_2d_dg.f90:1516-1555"""
        u = np.array(modes, dtype=np.float64, order="C", copy=True)
        _check(lib().wb_dg2d_apply_limiter(self._h, _ptr(u, self.shape)))
        return u

    def compute_max_speed(self, mean_mode):
        """compute_max_speed(u(:,:,:,1,1),...)  2d/benchmark_2d_dg.f90:826-870 -> (cs_max, v_xmax, v_ymax, speed_max)"""
        a = [C.c_double() for _ in range(4)]
        mm = np.ascontiguousarray(mean_mode)
        _check(lib().wb_dg2d_compute_max_speed(self._h, _ptr(mm), *[C.byref(v) for v in a]))
        return tuple(v.value for v in a)

    def evolve(self, u_nodes, x=None, y=None, tend=1.0, max_iter=-1):
        """evolve(u,x,y,u_eq)  2d/benchmark_2d_dg.f90:624-775 -> (u_nodes_new, iters, t, last_dt)"""
        u = np.array(u_nodes, dtype=np.float64, order="C", copy=True)
        it = C.c_int(); t = C.c_double(); dt = C.c_double()
        _check(lib().wb_dg2d_evolve(self._h, _ptr(u, self.shape), _ptr(x, self.shape[:-1]) if x is not None else None, _ptr(y, self.shape[:-1]) if y is not None else None,
                                    C.c_double(tend), C.c_int(max_iter), C.byref(it), C.byref(t), C.byref(dt)))
        return u, it.value, t.value, dt.value

    def upload(self, u_nodes, x=None, y=None):
        _check(lib().wb_dg2d_upload(self._h, _ptr(u_nodes, self.shape), _ptr(x, self.shape[:-1]) if x is not None else None, _ptr(y, self.shape[:-1]) if y is not None else None))

    def init_device(self, ninit, eta=F32(0.1)):
        _check(lib().wb_dg2d_init_device(self._h, C.c_int(ninit), C.c_double(eta)))

    def compute_error_resident(self, ninit, shift_x=0.0, shift_y=0.0, eta=F32(0.1)):
        """compute_error of the resident state against initial condition `ninit` translated by (shift_x, shift_y), on the device
        -> (lmax[4], l1[4], l2[4] before the sqrt)"""
        a = np.zeros(4); b = np.zeros(4); c = np.zeros(4)
        _check(lib().wb_dg2d_compute_error_resident(self._h, C.c_int(ninit), C.c_double(eta), C.c_double(shift_x), C.c_double(shift_y),
                                                    _ptr(a), _ptr(b), _ptr(c)))
        return a, b, c

    def output_file(self, path, var=1, nequilibrium=3, wait=True):
        """output_file(x,y,nodes,var,filen)  2d/benchmark_2d_dg.f90:468-495 of the resident state (asynchronous writer)"""
        _check(lib().wb_dg2d_output_file(self._h, C.c_int(var), C.c_int(nequilibrium), C.c_char_p(os.fsencode(path))))
        if wait:
            self.output_wait()

    def output_wait(self):
        _check(lib().wb_dg2d_output_wait(self._h))

    def get_initial_conditions(self, ninit, eta=F32(0.1)):
        """get_initial_conditions(x,y,u,...)  2d/benchmark_2d_dg.f90:122-466, ninit 1..12 -> nodal conserved state (owned rows)"""
        out = np.empty(self.shape)
        _check(lib().wb_dg2d_get_initial_conditions(self._h, C.c_int(ninit), C.c_double(eta), _ptr(out)))
        return out

    def step_async(self, nsteps, tend=1e300):
        _check(lib().wb_dg2d_step_async(self._h, C.c_int(nsteps), C.c_double(tend)))

    def sync(self):
        it = C.c_int(); t = C.c_double(); dt = C.c_double()
        _check(lib().wb_dg2d_sync(self._h, C.byref(it), C.byref(t), C.byref(dt)))
        return it.value, t.value, dt.value

    def download(self):
        out = np.empty(self.shape)
        _check(lib().wb_dg2d_download(self._h, _ptr(out)))
        return out

    def download_modes(self):
        out = np.empty(self.shape)
        _check(lib().wb_dg2d_download_modes(self._h, _ptr(out)))
        return out


# ================================================================================================== 1D FV
class FVM1DParams(C.Structure):
    """wb_fvm1d_params; defaults follow fvm_commons.f90."""
    _fields_ = [("nx", C.c_int), ("nvar", C.c_int), ("bc", C.c_int), ("source", C.c_int), ("n", C.c_int),
                ("gamma", C.c_double), ("boxlen", C.c_double), ("device", C.c_int)]


class _Handle:
    _destroy = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            getattr(lib(), self._destroy)(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class FVM1D(_Handle):
    """fvm.f90: plain 1D finite volumes.  u(nvar,nx) == numpy (nx, 3)."""
    _destroy = "wb_fvm1d_destroy"

    def __init__(self, nx=200, bc=2, source=2, n=3, gamma=F32(1.4), boxlen=1.0, device=-1):
        self.params = FVM1DParams(nx, 3, bc, source, n, gamma, boxlen, device)
        self._h = C.c_void_p()
        _check(lib().wb_fvm1d_create(C.byref(self._h), C.byref(self.params)))
        self.shape = (nx, 3)

    def compute_update(self, u):
        """compute_update(u,dudt)  fvm.f90:188-251"""
        d = np.empty(self.shape)
        _check(lib().wb_fvm1d_compute_update(self._h, _ptr(u), _ptr(d)))
        return d

    def compute_max_speed(self, u):
        c = C.c_double()
        _check(lib().wb_fvm1d_compute_max_speed(self._h, _ptr(u), C.byref(c)))
        return c.value

    def evolve(self, u, tend, max_iter=-1):
        u = np.array(u, dtype=np.float64, order="C", copy=True)
        it = C.c_int(); t = C.c_double(); dt = C.c_double()
        _check(lib().wb_fvm1d_evolve(self._h, _ptr(u), C.c_double(tend), C.c_int(max_iter), C.byref(it), C.byref(t), C.byref(dt)))
        return u, it.value, t.value, dt.value


class FV1DParams(C.Structure):
    """wb_fv1d_params; defaults follow parameters.f90."""
    _fields_ = [("nx", C.c_int), ("nvar", C.c_int), ("bc", C.c_int), ("nequilibrium", C.c_int), ("solver", C.c_int),
                ("gamma", C.c_double), ("boxlen", C.c_double), ("device", C.c_int)]


SOLVERS_1D = {"FVM": 1, "EQL": 2, "WB1": 3}


class FV1D(_Handle):
    """benchmark_1d.f90: 'FVM' | 'EQL' | 'WB1'.  u(nvar,nx) == numpy (nx, 3)."""
    _destroy = "wb_fv1d_destroy"

    def __init__(self, nx=128, bc=2, nequilibrium=2, solver="WB1", gamma=F32(1.4), boxlen=1.0, device=-1):
        self.params = FV1DParams(nx, 3, bc, nequilibrium, SOLVERS_1D[solver], gamma, boxlen, device)
        self._h = C.c_void_p()
        _check(lib().wb_fv1d_create(C.byref(self._h), C.byref(self.params)))
        self.shape = (nx, 3)

    def _upd(self, fn, u, w_eq):
        d = np.empty(self.shape)
        _check(getattr(lib(), fn)(self._h, _ptr(u), _ptr(w_eq) if w_eq is not None else None, _ptr(d)))
        return d

    def compute_update(self, u, w_eq):
        """compute_update ('EQL')  benchmark_1d.f90:263-377"""
        return self._upd("wb_fv1d_compute_update", u, w_eq)

    def compute_update_fvm(self, u, w_eq):
        """compute_update_fvm ('FVM')  benchmark_1d.f90:454-549"""
        return self._upd("wb_fv1d_compute_update_fvm", u, w_eq)

    def compute_update_sr(self, u, w_eq=None):
        """compute_update_sr ('WB1')  benchmark_1d.f90:553-747"""
        return self._upd("wb_fv1d_compute_update_sr", u, w_eq)

    def compute_max_speed(self, u):
        c = C.c_double()
        _check(lib().wb_fv1d_compute_max_speed(self._h, _ptr(u), C.byref(c)))
        return c.value

    def evolve(self, u, w_eq, tend, max_iter=-1):
        """evolve(u,u_eq,x)  benchmark_1d.f90:200-261"""
        u = np.array(u, dtype=np.float64, order="C", copy=True)
        it = C.c_int(); t = C.c_double(); dt = C.c_double()
        _check(lib().wb_fv1d_evolve(self._h, _ptr(u), _ptr(w_eq), C.c_double(tend), C.c_int(max_iter), C.byref(it),
                                    C.byref(t), C.byref(dt)))
        return u, it.value, t.value, dt.value


# ================================================================================================== 1D DG
class DG1DParams(C.Structure):
    """wb_dg1d_params; defaults follow dg_commons.f90."""
    _fields_ = [("n", C.c_int), ("nx", C.c_int), ("nvar", C.c_int), ("riemann", C.c_int), ("source", C.c_int),
                ("gamma", C.c_double), ("boxlen", C.c_double), ("device", C.c_int), ("bc", C.c_int), ("use_limiter", C.c_int)]


class DG1D(_Handle):
    """dg_with_source.f90, integrator 'RKi' (perturbation form).  u(nvar,n,nx) == numpy (nx, n, 3)."""
    _destroy = "wb_dg1d_destroy"

    def __init__(self, n=3, nx=128, riemann=2, source=2, gamma=F32(1.4), boxlen=1.0, device=-1, bc=5, use_limiter=False):
        self.params = DG1DParams(n, nx, 3, riemann, source, gamma, boxlen, device, bc, int(use_limiter))
        self._h = C.c_void_p()
        _check(lib().wb_dg1d_create(C.byref(self._h), C.byref(self.params)))
        self.shape = (nx, n, 3)

    def quadrature(self):
        x = np.zeros(self.params.n); w = np.zeros(self.params.n)
        _check(lib().wb_dg1d_quadrature(self._h, _ptr(x), _ptr(w)))
        return x, w

    def compute_update_exact_delta(self, delta_u, u_eq):
        """compute_update_exact_delta(delta_u,u_eq,dudt)  dg_with_source.f90:1749-2031"""
        d = np.empty(self.shape)
        _check(lib().wb_dg1d_compute_update_exact_delta(self._h, _ptr(delta_u), _ptr(u_eq), _ptr(d)))
        return d

    def compute_max_speed(self, u_nodes):
        c = C.c_double()
        _check(lib().wb_dg1d_compute_max_speed(self._h, _ptr(u_nodes), C.byref(c)))
        return c.value

    def evolve(self, delta_u, u_eq, uinit, tend, max_iter=-1):
        """main loop, integrator 'RKi'  dg_with_source.f90:173-336 -> (delta_u, uinit, iters, t, last_dt)"""
        d = np.array(delta_u, dtype=np.float64, order="C", copy=True)
        ui = np.array(uinit, dtype=np.float64, order="C", copy=True)
        it = C.c_int(); t = C.c_double(); dt = C.c_double()
        _check(lib().wb_dg1d_evolve(self._h, _ptr(d), _ptr(u_eq), _ptr(ui), C.c_double(tend), C.c_int(max_iter), C.byref(it),
                                    C.byref(t), C.byref(dt)))
        return d, ui, it.value, t.value, dt.value

    def compute_update(self, u):
        """compute_update(u,dudt)  dg_with_source.f90:807-1028"""
        d = np.empty(self.shape)
        _check(lib().wb_dg1d_compute_update(self._h, _ptr(u), _ptr(d)))
        return d

    def limiter(self, u):
        """limiter(u)  dg_with_source.f90:414-519"""
        v = np.array(u, dtype=np.float64, order="C", copy=True)
        _check(lib().wb_dg1d_limiter(self._h, _ptr(v)))
        return v

    def evolve_rk(self, integrator, u, delta_u, u_eq, uinit, tend, max_iter=-1):
        """main loop with 'RK1' | 'RK2' | 'RK3' | 'RK4'  dg_with_source.f90:173-227 -> (u, uinit, iters, t, last_dt)"""
        uu = np.array(u, dtype=np.float64, order="C", copy=True)
        ui = np.array(uinit, dtype=np.float64, order="C", copy=True)
        it = C.c_int(); t = C.c_double(); dt = C.c_double()
        _check(lib().wb_dg1d_evolve_rk(self._h, C.c_int({"RK1": 1, "RK2": 2, "RK3": 3, "RK4": 4}[integrator]), _ptr(uu), _ptr(delta_u),
                                       _ptr(u_eq), _ptr(ui), C.c_double(tend), C.c_int(max_iter), C.byref(it), C.byref(t), C.byref(dt)))
        return uu, ui, it.value, t.value, dt.value

    def compute_update_exact(self, u, u_eq_modes):
        """compute_update_exact(u,u_eq_modes,dudt)  dg_with_source.f90:1380-1744 (bc 4 | 5)"""
        d = np.empty(self.shape)
        _check(lib().wb_dg1d_compute_update_exact(self._h, _ptr(u), _ptr(u_eq_modes), _ptr(d)))
        return d

    def limiter_TDV(self, u):
        """limiter_TDV(u)  dg_with_source.f90:520-600 (use_limiter = .false.)"""
        v = np.array(u, dtype=np.float64, order="C", copy=True)
        _check(lib().wb_dg1d_limiter_tdv(self._h, _ptr(v)))
        return v

    def limiter_cons(self, u):
        """limiter_cons(u)  dg_with_source.f90:602-734"""
        v = np.array(u, dtype=np.float64, order="C", copy=True)
        _check(lib().wb_dg1d_limiter_cons(self._h, _ptr(v)))
        return v

    def evolve_w(self, integrator, u, delta_u, u_eq_nodes, u_eq_modes, uinit, tend, max_iter=-1):
        """main loop with 'RKw' | 'RKe'  dg_with_source.f90:229-280 -> (u, delta_u, uinit, iters, t, last_dt)"""
        uu = np.array(u, dtype=np.float64, order="C", copy=True)
        dd = np.array(delta_u, dtype=np.float64, order="C", copy=True)
        ui = np.array(uinit, dtype=np.float64, order="C", copy=True)
        it = C.c_int(); t = C.c_double(); dt = C.c_double()
        _check(lib().wb_dg1d_evolve_w(self._h, C.c_int({"RKw": 5, "RKe": 6}[integrator]), _ptr(uu), _ptr(dd), _ptr(u_eq_nodes),
                                      _ptr(u_eq_modes), _ptr(ui), C.c_double(tend), C.c_int(max_iter), C.byref(it), C.byref(t),
                                      C.byref(dt)))
        return uu, dd, ui, it.value, t.value, dt.value
