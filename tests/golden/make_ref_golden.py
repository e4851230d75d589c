"""Golden vectors produced by THE REFERENCE'S OWN SOURCE TEXT.

The image has no Fortran compiler, so instead of building the reference this script executes its .f90 files, unmodified,
where they lie under /root/reference, with the small Fortran-90 interpreter in oracle/f90interp.py (semantics of
gfortran on x86-64 without FMA contraction: real(4) literals, mixed-mode promotion, sequence association, libm
transcendentals).  Sizes and switches are set the way the reference is configured -- by giving the `parameter`
constants of its parameter modules other values (Interp.override; the files are not touched).

Inputs and outputs of the hot-path routines are stored in the ORACLE's array layout (C order = the Fortran array
transposed) under tests/golden/ref_*.npz; tests/test_reference_pins.py checks the C oracle (oracle/*.c) against
them, and the GPU tests check the CUDA path against some of them directly.  /root/reference is only needed to
RE-GENERATE the vectors:      python tests/golden/make_ref_golden.py [fv2d dg2d fv1d dg1d]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.f90interp import FortranBoundsError, Interp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WB_REFERENCE", "/root/reference")


def F(*shape):
    return np.zeros(shape, order="F")


def C(a):
    """Fortran array -> the oracle's C-order array (same bytes)."""
    return np.array(np.asarray(a).T, order="C", copy=True)


def scalar(v=0.0):
    return np.array(v, dtype=np.float64)


def note_calls(out, tag, it):
    out[f"{tag}/calls"] = np.array(sorted(f"{k}:{v}" for k, v in it.calls.items()))
    it.calls.clear()


# ------------------------------------------------------------------------------------------------ 2D FV
def fv2d_interp(**kv):
    it = Interp()
    it.load(f"{REF}/parameters_2d.f90").load(f"{REF}/benchmark_2d.f90")
    it.override("parameters_2d", **kv)
    return it


def fv2d():
    out = {}
    rng = np.random.default_rng(20161017)
    cases = [  # tag, nx, ny, ninit, nequilibrium, steps, random perturbation
        ("pert", 12, 12, 3, 2, 3, 0.0), ("ragged", 13, 9, 3, 2, 2, 0.0), ("riemann", 12, 10, 4, 2, 3, 0.0),
        ("eq1", 8, 10, 1, 1, 2, 0.0), ("hydro", 10, 10, 2, 2, 2, 0.0), ("random", 9, 11, 3, 2, 2, 0.1),
        ("eq3_rand", 10, 8, 2, 3, 1, 0.05),
        # the headline workload (BASELINE config 3: ninit 3, nequilibrium 2) on larger grids, more steps
        ("headline_32", 32, 32, 3, 2, 6, 0.0), ("headline_48x40_rand", 48, 40, 3, 2, 4, 0.02),
    ]
    for tag, nx, ny, ninit, neq, steps, amp in cases:
        it = fv2d_interp(nx=nx, ny=ny, ninit=ninit, nequilibrium=neq)
        x, y = F(nx, ny), F(nx, ny)
        it.call("get_coords", x, y, nx, ny)
        u, weq = F(4, nx, ny), F(4, nx, ny)
        it.call("get_initial_conditions", x, y, u, nx, ny)
        it.call("get_equilibrium_solution", x, y, weq, nx, ny)
        out[f"{tag}/u_ic"] = C(u)
        if amp:
            u *= 1.0 + amp * rng.standard_normal(u.shape)
            u[1] = amp * 3 * rng.standard_normal((nx, ny))
            u[2] = amp * 2 * rng.standard_normal((nx, ny))
        w = F(4, nx, ny)
        it.call("compute_primitive", u, w, nx, ny)
        u_back = F(4, nx, ny)
        it.call("compute_conservative", w, u_back, nx, ny)
        cmax = scalar()
        it.call("compute_max_speed", u, cmax)
        dudt, dplain = F(4, nx, ny), F(4, nx, ny)
        it.call("compute_update_exact", u, weq, dudt)
        try:
            it.call("compute_update", u, weq, dplain)
            plain_ok = True
        except FortranBoundsError:       # benchmark_2d.f90:418 `iright = ny` indexes x with ny: out of bounds when nx < ny
            assert nx < ny
            plain_ok = False
        out[f"{tag}/meta"] = np.array([nx, ny, ninit, neq, steps])
        for k, v in (("x", x), ("y", y), ("u", u), ("weq", weq), ("w", w), ("u_back", u_back), ("dudt", dudt), ("dudt_plain", dplain)):
            if k != "dudt_plain" or plain_ok:
                out[f"{tag}/{k}"] = C(v)
        out[f"{tag}/cmax"] = cmax.copy()
        # evolve: reproduce the clock of `steps` steps by choosing tend just below the time reached after them
        # (the reference has no iteration limit; `tend` is a module variable)
        un = np.array(u, order="F", copy=True)
        # first find the dt sequence: tend tiny -> 1 step; we simply run with a huge step budget bounded by tend
        t_end = 0.0
        probe = np.array(u, order="F", copy=True)
        it.override("parameters_2d", nx=nx, ny=ny, ninit=ninit, nequilibrium=neq, tend=1e-300)
        dts = []
        for _ in range(steps):
            c = scalar()
            it.call("compute_max_speed", probe, c)
            dx = 1.0 / nx
            dts.append(0.5 * dx / float(c) * 0.5)
            it.call("evolve", probe, weq)       # tend=1e-300: exactly one step
        t_end = sum(dts[:-1]) + 0.5 * dts[-1]   # the loop `do while (t<tend)` then takes exactly `steps` steps
        it.override("parameters_2d", nx=nx, ny=ny, ninit=ninit, nequilibrium=neq, tend=t_end)
        it.call("evolve", un, weq)
        assert np.array_equal(un, probe), "one-step-at-a-time evolve differs from the tend-bounded one"
        out[f"{tag}/u_evolved"] = C(un)
        out[f"{tag}/tend"] = np.array(t_end)
        note_calls(out, tag, it)
        print(f"fv2d {tag}: max|dudt| = {np.abs(dudt).max():.3e}, steps = {steps}", flush=True)
    np.savez_compressed(os.path.join(HERE, "ref_fv2d.npz"), **out)


# ------------------------------------------------------------------------------------------------ 2D DG
def dg2d_interp(**kv):
    it = Interp()
    for f in ("parameters_dg_2d.f90", "legendre.f90", "limiters.f90", "benchmark_2d_dg.f90"):
        it.load(f"{REF}/2d/{f}")
    it.override("parameters_dg_2d", **kv)
    return it


DG2D_CASES = [  # tag, n (nx=ny), m (mx=my), bc, source, grad_phi_case, flux_type, limiter_type, solver, ninit, steps
    ("pulse_o3_onp", 3, 3, 1, 1, 2, "llf1", "ONP", "RK4", 1, 2),
    ("pulse_o2_onp", 4, 2, 1, 1, 2, "llf1", "ONP", "RK4", 1, 2),
    ("pulse_o1", 4, 1, 1, 1, 2, "llf1", "ONP", "RK4", 1, 2),
    ("shipped_flux", 3, 2, 1, 1, 2, "llf", "ONP", "RK4", 1, 1),
    ("hydro_o3_g1", 3, 3, 2, 2, 1, "llf1", "ONP", "RK4", 2, 1),
    ("hydro_o2_kepler", 4, 2, 2, 2, 2, "llf1", "ONP", "EQL", 2, 2),
    ("riemann_o3_hio", 4, 3, 2, 1, 2, "llf1", "HIO", "RK4", 3, 1),
    ("riemann_o2_1or", 4, 2, 2, 1, 2, "llf1", "1OR", "EQL", 4, 2),
    ("riemann_o3_low", 3, 3, 2, 1, 2, "llf1", "LOW", "DEB", 4, 2),
    ("advsink_o2", 3, 2, 1, 3, 2, "llf1", "ONP", "SS4", 1, 1),
    ("riemann_o4_onp", 3, 4, 3, 1, 2, "llf1", "ONP", "RK4", 5, 1),
    ("pulse_o3_hll2", 3, 3, 1, 1, 2, "hll2", "ONP", "RK4", 1, 1),
    ("riemann_o2_hllc", 4, 2, 2, 1, 2, "hllc", "ONP", "RK4", 3, 1),
    ("riemann_o3_pos", 4, 3, 2, 1, 2, "llf1", "POS", "EQL", 3, 2),
    # rotating disk in a 6 x 6 box (boxlen is a module variable): Keplerian grad_phi with the softened core and
    # special_boundary_conditions (:1481-1514), which freezes the update outside r = 2
    ("disk_o2_ninit12", 6, 2, 2, 2, 2, "llf1", "ONP", "RK4", 12, 1),
    # the headline workload (BASELINE config 4: order 3, llf1, ONP, SSPRK(5,4), periodic pulse) on a larger grid, 3 steps
    ("headline_o3_n6", 6, 3, 1, 1, 2, "llf1", "ONP", "RK4", 1, 3),
]


def dg2d(only=None):
    out = {}
    path = os.path.join(HERE, "ref_dg2d.npz")
    if only and os.path.exists(path):
        out = dict(np.load(path))
    for tag, n, m, bc, source, gcase, flux, lim, solver, ninit, steps in DG2D_CASES:
        if only and tag not in only:
            continue
        t0 = time.time()
        kv = dict(nx=n, ny=n, mx=m, my=m, bc=bc, source=source, grad_phi_case=gcase, flux_type=flux, limiter_type=lim,
                  solver=solver, ninit=ninit)
        box = 6.0 if ninit == 12 else 1.0
        if box != 1.0:
            kv.update(boxlen_x=box, boxlen_y=box)
        it = dg2d_interp(**kv)
        x, y = F(n, n, m, m), F(n, n, m, m)
        it.call("get_coords", x, y, n, n, m, m)
        nodes = F(4, n, n, m, m)
        it.call("get_initial_conditions", x, y, nodes, n, n, m, m)
        modes = F(4, n, n, m, m)
        it.call("get_modes_from_nodes", nodes, modes, n, n, m, m)
        back = F(4, n, n, m, m)
        it.call("get_nodes_from_modes", modes, back, n, n, m, m)
        ueq = F(4, n, n, m, m)
        dudt = F(4, n, n, m, m)
        it.call("compute_update", modes, x, y, ueq, dudt)
        lim_modes = np.array(modes, order="F", copy=True)
        it.call("apply_limiter", lim_modes)
        # a rougher state for the limiter: the update direction added with a large step
        rough = np.asfortranarray(modes + 0.05 * dudt / max(np.abs(dudt).max(), 1e-300) * np.abs(modes[0]).max())
        rough_in = rough.copy(order="F")
        it.call("apply_limiter", rough)
        sp = [scalar() for _ in range(4)]
        mean = np.asfortranarray(modes[:, :, :, 0, 0])
        it.call("compute_max_speed", mean, *sp)
        out[f"{tag}/meta"] = np.array([n, m, bc, source, gcase, ninit, steps])
        out[f"{tag}/names"] = np.array([flux, lim, solver])
        out[f"{tag}/boxlen"] = np.array(box)
        for k, v in (("x", x), ("y", y), ("nodes", nodes), ("modes", modes), ("nodes_back", back), ("dudt", dudt),
                     ("limited", lim_modes), ("rough_in", rough_in), ("rough_limited", rough)):
            out[f"{tag}/{k}"] = C(v)
        out[f"{tag}/speeds"] = np.array([float(s) for s in sp])
        # evolve: the reference loop has no iteration limit; bound it with `tend` (module variable).  dt of the first step:
        cs, vx, vy, _ = (float(s) for s in sp)
        # the limiter runs before the first step and can change the mean-mode speeds only through the modes it scales
        # (never the means), so the first dt is known; later ones are not needed: tend = (steps - 0.5) * dt0 gives
        # `steps` steps unless dt grows by more than 2x, and the last step is clipped by min(tend - t, ...).
        gll = (2 * (m - 1) + 3) // 2
        gll_w_1 = 1.0 if m == 1 else 1.0 / (float(gll * (gll - 1)) + float(np.float32(1e-10)))
        dx = box / n
        cfl = float(np.float32(0.2))
        dt0 = cfl * min(1.0 / 9.0, gll_w_1 / 2.0) / ((abs(vx) + cs) / dx + (abs(vy) + cs) / dx)
        tend = (steps - 0.5) * dt0
        it.override("parameters_dg_2d", tend=tend, **kv)
        un = np.array(nodes, order="F", copy=True)
        it.call("evolve", un, x, y, ueq)
        out[f"{tag}/tend"] = np.array(tend)
        out[f"{tag}/nodes_evolved"] = C(un)
        out[f"{tag}/updates_in_evolve"] = np.array(it.calls.get("compute_update", 0))
        note_calls(out, tag, it)
        print(f"dg2d {tag}: {time.time() - t0:.1f} s, max|dudt| = {np.abs(dudt).max():.3e}, limiter changed "
              f"{np.abs(rough - rough_in).max():.2e}, evolve changed {np.abs(un - nodes).max():.2e}", flush=True)
        np.savez_compressed(path, **out)


def dg2d_limiters():
    """apply_limiter on rough data: every limiter has to act (negative density / pressure points, steep linear modes)."""
    out = {}
    rng = np.random.default_rng(99)
    for lim in ("ONP", "HIO", "1OR", "LOW", "POS"):
        for n, m, bc in ((3, 3, 1), (4, 2, 2), (3, 2, 3)):
            tag = f"{lim.lower()}_n{n}_m{m}_bc{bc}"
            t0 = time.time()
            it = dg2d_interp(nx=n, ny=n, mx=m, my=m, bc=bc, limiter_type=lim, flux_type="llf1", ninit=1)
            modes = F(4, n, n, m, m)
            modes[0, :, :, 0, 0] = 1.0 + 0.5 * rng.random((n, n))
            modes[1, :, :, 0, 0] = 0.3 * rng.standard_normal((n, n))
            modes[2, :, :, 0, 0] = 0.3 * rng.standard_normal((n, n))
            modes[3, :, :, 0, 0] = 2.5 + rng.random((n, n))
            hi = 0.08 * rng.standard_normal((4, n, n, m, m))
            hi[:, :, :, 0, 0] = 0.0
            modes += hi
            modes[0, 0, 1, 0, 1] = 0.9            # density dips below zero at a face point of cell (1,2)
            modes[3, 1, 0, 1, 0] = -2.0           # energy (pressure) dips in cell (2,1)
            if m > 2:
                modes[0, 2, 2, 2, 2] = 0.4
            inp = modes.copy(order="F")
            it.call("apply_limiter", modes)
            out[f"{tag}/meta"] = np.array([n, m, bc])
            out[f"{tag}/limiter"] = np.array(lim)
            out[f"{tag}/in"] = C(inp)
            out[f"{tag}/out"] = C(modes)
            note_calls(out, tag, it)
            print(f"dg2d limiter {tag}: {time.time() - t0:.1f} s, changed {np.abs(modes - inp).max():.3e} in "
                  f"{int((np.abs(modes - inp).reshape(4, n * n, m * m).max(axis=(0, 2)) > 0).sum())}/{n * n} cells", flush=True)
            np.savez_compressed(os.path.join(HERE, "ref_dg2d_limiters.npz"), **out)


def dg2d_po3():
    """limiter_type 'PO3' (limiter_positivity_2, 2d/limiters.f90:1587-1711): characteristic-variable minmod + nodal reset,
    on rough data and on smooth data -> ref_dg2d_po3.npz"""
    out = {}
    rng = np.random.default_rng(1587)
    for n, m, bc, rough in ((3, 3, 1, True), (4, 2, 2, True), (4, 3, 1, False), (3, 4, 3, True), (5, 2, 1, False)):
        tag = f"po3_n{n}_m{m}_bc{bc}_{'rough' if rough else 'smooth'}"
        t0 = time.time()
        it = dg2d_interp(nx=n, ny=n, mx=m, my=m, bc=bc, limiter_type="PO3", flux_type="llf1", ninit=1)
        modes = F(4, n, n, m, m)
        modes[0, :, :, 0, 0] = 1.0 + 0.5 * rng.random((n, n))
        modes[1, :, :, 0, 0] = 0.3 * rng.standard_normal((n, n))
        modes[2, :, :, 0, 0] = 0.3 * rng.standard_normal((n, n))
        modes[3, :, :, 0, 0] = 2.5 + rng.random((n, n))
        hi = (0.08 if rough else 0.004) * rng.standard_normal((4, n, n, m, m))
        hi[:, :, :, 0, 0] = 0.0
        modes += hi
        if rough:
            modes[0, 0, 1, 0, 1] = 0.9            # density dips below zero at a node of cell (1,2)
            modes[3, 1, 0, 1, 0] = -2.0           # energy (pressure) dips in cell (2,1)
        inp = modes.copy(order="F")
        it.call("apply_limiter", modes)
        out[f"{tag}/meta"] = np.array([n, m, bc])
        out[f"{tag}/limiter"] = np.array("PO3")
        out[f"{tag}/in"] = C(inp)
        out[f"{tag}/out"] = C(modes)
        note_calls(out, tag, it)
        print(f"dg2d PO3 {tag}: {time.time() - t0:.1f} s, changed {np.abs(modes - inp).max():.3e}, nan {int(np.isnan(modes).sum())}", flush=True)
    np.savez_compressed(os.path.join(HERE, "ref_dg2d_po3.npz"), **out)


def dg2d_other_limiters():
    """The limiter_type branches of apply_limiter that are NOT built, and why -- observed by running them:
    'ROS' and 'KRI' index out of bounds / pass a rank-1 section to a rank-5 dummy (undefined behaviour in the reference),
    'COC' overwrites the nodal pressure with the literal 10e-5 (an abandoned experiment), 'PO3' returns NaN at order 3,
    '1DL' calls limiter_1d, which does not exist."""
    import json
    from oracle.f90interp import FortranError
    rng = np.random.default_rng(3)
    facts = {}
    for lim in ("ROS", "KRI", "COC", "PO3", "1DL"):
        for n, m, bc in ((3, 3, 1), (4, 2, 2)):
            it = dg2d_interp(nx=n, ny=n, mx=m, my=m, bc=bc, limiter_type=lim, flux_type="llf1", ninit=1)
            modes = F(4, n, n, m, m)
            modes[0, :, :, 0, 0] = 1 + 0.5 * rng.random((n, n)); modes[3, :, :, 0, 0] = 2.5 + rng.random((n, n))
            modes[1, :, :, 0, 0] = 0.3 * rng.standard_normal((n, n)); modes[2, :, :, 0, 0] = 0.3 * rng.standard_normal((n, n))
            hi = 0.08 * rng.standard_normal((4, n, n, m, m)); hi[:, :, :, 0, 0] = 0; modes += hi
            key = f"{lim}_n{n}_m{m}_bc{bc}"
            try:
                it.call("apply_limiter", modes)
                nodes = F(4, n, n, m, m); w = F(4, n, n, m, m)
                it.call("get_nodes_from_modes", modes, nodes, n, n, m, m)
                it.call("compute_primitive", nodes, w, n, n, m, m)
                facts[key] = {"status": "ran", "nan": int(np.isnan(modes).sum()),
                              "pressure_min": float(np.nanmin(w[3])) if not np.isnan(w[3]).all() else None,
                              "pressure_max": float(np.nanmax(w[3])) if not np.isnan(w[3]).all() else None}
            except FortranError as e:
                facts[key] = {"status": "error", "message": str(e)[:160]}
            print(key, facts[key], flush=True)
    json.dump(facts, open(os.path.join(HERE, "ref_dg2d_other_limiters.json"), "w"), indent=1)


def dg2d_error_norms():
    """compute_error (2d/benchmark_2d_dg.f90:23-89): L1 / L2 accumulators (before the sqrt) and the max errors it prints."""
    out = {}
    rng = np.random.default_rng(17)
    for tag, n, m, ninit in (("pulse_o3", 4, 3, 1), ("pulse_o2", 5, 2, 1), ("hydro_o3", 3, 3, 2)):
        it = dg2d_interp(nx=n, ny=n, mx=m, my=m, ninit=ninit, flux_type="llf1")
        x, y = F(n, n, m, m), F(n, n, m, m)
        it.call("get_coords", x, y, n, n, m, m)
        u0 = F(4, n, n, m, m)
        it.call("get_initial_conditions", x, y, u0, n, n, m, m)
        u = np.asfortranarray(u0 * (1 + 0.01 * rng.standard_normal(u0.shape)) + 0.003 * rng.standard_normal(u0.shape))
        fr = it.call("compute_error", u, x, y, np.array(0.3), F(4, n, n, m, m))
        out[f"{tag}/meta"] = np.array([n, m, ninit])
        out[f"{tag}/x"] = C(x); out[f"{tag}/y"] = C(y); out[f"{tag}/u"] = C(u); out[f"{tag}/u_init"] = C(fr["u_init"])
        out[f"{tag}/l1"] = np.array(fr["l1norm"]); out[f"{tag}/l2"] = np.array(fr["l2norm"])
        out[f"{tag}/lmax"] = np.array([np.abs(u[v] - u0[v]).max() for v in range(4)])
        print(f"dg2d compute_error {tag}: l1 {fr['l1norm']}, l2 {fr['l2norm']}", flush=True)
    np.savez_compressed(os.path.join(HERE, "ref_dg2d_error_norms.npz"), **out)


def test2d():
    """2d/test2d.f90 -- the reference's only test program -- run as shipped (nx = ny = 8, mx = my = 2, 1000 round trips
    modes <-> nodes through 2d/commons.f90); it prints maxval(u - nodes) and minval(u - nodes)."""
    it = Interp()
    for f in ("parameters_dg_2d.f90", "legendre.f90", "commons.f90", "test2d.f90"):
        it.load(f"{REF}/2d/{f}")
    t0 = time.time()
    fr = it.run_program("dg")
    out = {"u": C(fr["u"]), "nodes": C(fr["nodes"]), "modes": C(fr["modes"]),
           "printed": np.array([np.max(fr["u"] - fr["nodes"]), np.min(fr["u"] - fr["nodes"])]),
           "calls": np.array(sorted(f"{k}:{v}" for k, v in it.calls.items()))}
    print(f"test2d: {time.time() - t0:.0f} s, max diff {out['printed'][0]:.3e}, min diff {out['printed'][1]:.3e}", flush=True)
    np.savez_compressed(os.path.join(HERE, "ref_test2d.npz"), **out)


# ------------------------------------------------------------------------------------------------ 1D FV
def fv1d():
    out = {}
    # fvm.f90 (program fvm: condinit, compute_update, compute_max_speed, RK2 main loop)
    for tag, nx, bc, source, ninit, steps in (("fvm_sod", 60, 2, 2, 4, 4), ("fvm_sine_periodic", 32, 1, 1, 1, 4),
                                              ("fvm_hydro", 40, 2, 2, 7, 3), ("fvm_ninit2_bc1_src2", 24, 1, 2, 2, 3),
                                              ("fvm_ninit3", 24, 2, 1, 3, 2), ("fvm_ninit5", 30, 2, 2, 5, 2),
                                              ("fvm_ninit6", 30, 2, 2, 6, 2)):
        it = Interp().load(f"{REF}/fvm_commons.f90").load(f"{REF}/fvm.f90")
        kv = dict(nx=nx, bc=bc, source=source, ninit=ninit)
        it.override("fvm_commons", tend=0.0, **kv)
        u0 = np.array(it.run_program("fvm")["u"], order="F", copy=True)       # tend = 0: the loop body never runs
        dudt = F(3, nx)
        it.call("compute_update", u0, dudt)
        c = scalar()
        it.call("compute_max_speed", u0, c)
        dt0 = float(np.float32(0.8)) * (1.0 / nx) / float(c) / 7.0
        tend = (steps - 0.5) * dt0
        it.override("fvm_commons", tend=tend, **kv)
        fr = it.run_program("fvm")
        out[f"{tag}/meta"] = np.array([0, nx, bc, source, ninit, int(fr["iter"])])
        out[f"{tag}/u0"] = C(u0); out[f"{tag}/dudt"] = C(dudt); out[f"{tag}/cmax"] = c.copy()
        out[f"{tag}/tend"] = np.array(tend); out[f"{tag}/un"] = C(fr["u"]); out[f"{tag}/clock"] = np.array([float(fr["t"]), float(fr["dt"])])
        note_calls(out, tag, it)
        print(f"fv1d {tag}: {int(fr['iter'])} steps, max|dudt| = {np.abs(dudt).max():.3e}", flush=True)
    # benchmark_1d.f90 ('FVM', 'EQL', 'WB1')
    rng = np.random.default_rng(7)
    for tag, nx, bc, neq, solver, ninit, eta, steps, amp in (
            ("b1_wb1_default", 64, 2, 2, "WB1", 2, None, 4, 0.0), ("b1_wb1_isentropic", 48, 2, 3, "WB1", 3, 0.0, 3, 0.0),
            ("b1_eql_bump", 48, 2, 2, "EQL", 2, 1e-3, 4, 0.0), ("b1_eql_bc3", 40, 3, 2, "EQL", 2, 1e-3, 3, 0.0),
            ("b1_fvm_bump", 37, 1, 2, "FVM", 2, 1e-3, 3, 0.0), ("b1_wb1_random", 32, 2, 2, "WB1", 2, 1e-2, 3, 0.05),
            ("b1_eql_random_bc1", 32, 1, 2, "EQL", 2, 1e-2, 2, 0.05), ("b1_fvm_bc3", 30, 3, 1, "FVM", 1, 0.0, 2, 0.02),
            ("b1_wb1_neq1", 32, 2, 1, "WB1", 1, 0.0, 2, 0.0)):
        it = Interp().load(f"{REF}/parameters.f90").load(f"{REF}/benchmark_1d.f90")
        kv = dict(nx=nx, bc=bc, nequilibrium=neq, solver=solver, ninit=ninit)
        if eta is not None:
            kv["eta"] = eta
        it.override("parameters", **kv)
        x = F(nx); u = F(3, nx); weq = F(3, nx)
        it.call("get_x", x, nx)
        it.call("get_initial_conditions", x, u, nx)
        it.call("get_equilibrium_solution", x, weq, nx)
        out[f"{tag}/u_ic"] = C(u)
        if amp:
            u *= 1.0 + amp * rng.standard_normal(u.shape)
            u[1] = amp * rng.standard_normal(nx)
        d = {}
        for name in ("compute_update", "compute_update_fvm", "compute_update_sr"):
            d[name] = F(3, nx)
            it.call(name, u, weq, d[name])
        c = scalar()
        it.call("compute_max_speed", u, c)
        dt0 = float(np.float32(0.8)) * (1.0 / nx) / float(c) / 3.0
        tend = (steps - 0.5) * dt0
        it.override("parameters", tend=tend, **kv)
        un = np.array(u, order="F", copy=True)
        fr = it.call("evolve", un, weq, x)
        out[f"{tag}/meta"] = np.array([1, nx, bc, neq, ninit, int(fr["iter"])])
        out[f"{tag}/solver"] = np.array(solver)
        out[f"{tag}/eta"] = np.array(float(it.get("parameters", "eta")))
        for k, v in (("x", x), ("u", u), ("weq", weq), ("dudt_eql", d["compute_update"]), ("dudt_fvm", d["compute_update_fvm"]),
                     ("dudt_sr", d["compute_update_sr"]), ("un", un)):
            out[f"{tag}/{k}"] = C(v)
        out[f"{tag}/cmax"] = c.copy(); out[f"{tag}/tend"] = np.array(tend)
        out[f"{tag}/clock"] = np.array([float(fr["t"]), float(fr["dt"])])
        note_calls(out, tag, it)
        print(f"fv1d {tag}: {int(fr['iter'])} steps, max|dudt_sr| = {np.abs(d['compute_update_sr']).max():.3e}", flush=True)
    np.savez_compressed(os.path.join(HERE, "ref_fv1d.npz"), **out)


# ------------------------------------------------------------------------------------------------ 1D DG
DG1D_CASES = [  # tag, n, nx, riemann, source, ninit, pert, bc, use_limiter, integrator, steps
    ("rki_default_hllc", 3, 32, 2, 2, 8, None, 5, False, "RKi", 2),
    ("rki_llf_o2", 2, 24, 1, 2, 8, 1e-3, 5, False, "RKi", 2),
    ("rki_hllc_o1", 1, 20, 2, 2, 8, 1e-2, 5, False, "RKi", 3),
    ("rki_sod_nosource", 3, 20, 2, 1, 4, 0.0, 5, False, "RKi", 2),
    ("rki_steady_hllc", 3, 16, 2, 2, 7, 0.0, 5, False, "RKi", 2),
    ("rk2_periodic_lim", 3, 24, 2, 2, 8, 1e-2, 1, True, "RK2", 2),
    ("rk3_zerograd", 3, 20, 1, 2, 8, 1e-3, 2, True, "RK3", 2),
    ("rk4_reflect", 2, 17, 2, 2, 8, 1e-3, 3, True, "RK4", 2),
    ("rk1_sod_bc2", 3, 20, 2, 1, 4, 0.0, 2, False, "RK1", 3),
    ("rk3_sod_lim", 3, 24, 2, 1, 4, 0.0, 2, True, "RK3", 3),
    ("rk2_bc4", 2, 16, 1, 2, 8, 1e-2, 4, False, "RK2", 2),
    ("rkw_bc5", 3, 20, 2, 2, 8, 1e-3, 5, False, "RKw", 2),
    ("rkw_bc4_llf", 2, 16, 1, 2, 8, 1e-2, 4, False, "RKw", 2),
    ("rke_bc5", 3, 20, 1, 2, 8, 1e-3, 5, False, "RKe", 2),
    ("rke_lim_bc2", 3, 20, 2, 2, 8, 1e-2, 2, True, "RKe", 2),
]


def dg1d(only=None):
    out = {}
    path = os.path.join(HERE, "ref_dg1d.npz")
    if only and os.path.exists(path):
        out = dict(np.load(path))
    rng = np.random.default_rng(5)
    for tag, n, nx, riemann, source, ninit, pert, bc, use_limiter, integ, steps in DG1D_CASES:
        if only and tag not in only:
            continue
        t0 = time.time()
        it = Interp(oob="nan").load(f"{REF}/dg_commons.f90").load(f"{REF}/legendre.f90").load(f"{REF}/dg_with_source.f90")
        kv = dict(n=n, nx=nx, riemann=riemann, source=source, ninit=ninit, bc=bc, use_limiter=use_limiter, integrator=integ)
        if pert is not None:
            kv["pert"] = pert
        it.override("dg_commons", tend=0.0, **kv)
        # program dg projects the initial condition into `u` (:36-55) and then ZEROES `u` again in the set-up of delta_u
        # (:152), so as shipped every integrator but 'RKi'/'RKe' starts from u = 0 (0/0 = NaN).  The harness taps the
        # projection before it is zeroed (statement at line 62) and, for the timed run below, puts it back when the
        # clock is initialised (statement `t=0`, line 171): inputs are injected, no statement is changed or skipped.
        tap = {}
        it.probes = {62: lambda v: tap.setdefault("u", np.array(v["u"], order="F", copy=True))}
        fr = it.run_program("dg")              # tend = 0: set-up only (projections, equilibrium, module quadrature tables)
        it.probes = {}
        assert not np.any(fr["u"]), "the reference zeroes u before its main loop"
        u = tap["u"]
        du, ueq, q, ui = (np.array(fr[k], order="F", copy=True) for k in ("delta_u", "u_eq", "u_eq_modes", "u_init"))
        out[f"{tag}/meta"] = np.array([n, nx, riemann, source, ninit, bc, int(use_limiter), steps])
        out[f"{tag}/integrator"] = np.array(integ)
        out[f"{tag}/pert"] = np.array(float(it.get("dg_commons", "pert")))
        out[f"{tag}/quad"] = np.array([it.get("dg_commons", "chsi_quad"), it.get("dg_commons", "w_quad")])
        for k, v in (("u", u), ("du", du), ("ueq", ueq), ("q", q), ("ui", ui)):
            out[f"{tag}/{k}"] = C(v)
        c = scalar()
        it.call("compute_max_speed", ui, c)
        out[f"{tag}/cmax"] = c.copy()
        oob0 = it.oob_count
        d = F(3, n, nx)
        it.call("compute_update_exact_delta", du, ueq, d)
        out[f"{tag}/dudt_delta"] = C(d)
        out[f"{tag}/oob_delta"] = np.array(it.oob_count - oob0)
        # rougher modes for the limiters and the plain updates
        rough = np.array(u, order="F", copy=True)
        if n > 1:
            rough[:, 1:, :] += 1e-3 * rng.standard_normal(rough[:, 1:, :].shape)
            rough[0, 1, nx // 3] += 3.0 * rough[0, 0, nx // 3]          # a negative density trace
            rough[2, 1, 2 * nx // 3] += 30.0 * rough[2, 0, 2 * nx // 3]  # a negative energy trace
        out[f"{tag}/rough"] = C(rough)
        for name, key in (("compute_update", "dudt_plain"),):
            oob0 = it.oob_count
            d = F(3, n, nx)
            it.call(name, u, d)
            if it.oob_count == oob0:
                out[f"{tag}/{key}"] = C(d)
        oob0 = it.oob_count
        d = F(3, n, nx)
        it.call("compute_update_exact", u, q, d)
        if it.oob_count == oob0 or bc in (4, 5):
            out[f"{tag}/dudt_exact"] = C(d)
        out[f"{tag}/oob_exact"] = np.array(it.oob_count - oob0)
        for name, key in (("limiter", "lim"), ("limiter_cons", "lim_cons")) + ((("limiter_tdv", "lim_tdv"),) if not use_limiter else ()):
            v = np.array(rough, order="F", copy=True)
            oob0 = it.oob_count
            it.call(name, v)
            if it.oob_count == oob0:
                out[f"{tag}/{key}"] = C(v)
        dt0 = float(np.float32(0.9)) * (1.0 / nx) / float(c) / (2.0 * n + 1.0)
        tend = (steps - 0.5) * dt0
        it.override("dg_commons", tend=tend, **kv)

        def inject(v, u=u):
            v["u"][...] = u
        it.probes = {171: inject}
        fr = it.run_program("dg")
        it.probes = {}
        out[f"{tag}/tend"] = np.array(tend)
        out[f"{tag}/iters"] = np.array(int(fr["iter"]))
        out[f"{tag}/clock"] = np.array([float(fr["t"]), float(fr["dt"])])
        for k, v in (("u_end", fr["u"]), ("du_end", fr["delta_u"]), ("ureal_end", fr["ureal"])):
            out[f"{tag}/{k}"] = C(v)
        out[f"{tag}/oob_total"] = np.array(it.oob_count)
        note_calls(out, tag, it)
        print(f"dg1d {tag}: {time.time() - t0:.1f} s, {int(fr['iter'])} steps, oob reads {it.oob_count}, keys "
              f"{sorted(k.split('/')[1] for k in out if k.startswith(tag + '/') and 'dudt' in k or k.startswith(tag + '/lim'))}", flush=True)
        np.savez_compressed(path, **out)


# ------------------------------------------------------------------------------------------------ 2D DG initial conditions
DG2D_IC_CASES = [  # ninit, n (nx=ny), m (mx=my), boxlen
    (3, 4, 2, 1.0), (4, 4, 3, 1.0), (5, 5, 2, 1.0), (6, 4, 3, 10.0), (7, 8, 2, 6.0), (8, 6, 2, 1.0), (9, 6, 3, 1.0),
    (10, 4, 3, 1.0), (11, 8, 2, 6.0), (12, 6, 2, 6.0),
]


def dg2d_ics(only=None):
    """get_coords + get_initial_conditions (2d/benchmark_2d_dg.f90:93-120, :122-466) for ninit 3..12 -> ref_dg2d_ics.npz"""
    out = {}
    for ninit, n, m, box in DG2D_IC_CASES:
        t0 = time.time()
        kv = dict(nx=n, ny=n, mx=m, my=m, ninit=ninit)
        if box != 1.0:
            kv.update(boxlen_x=box, boxlen_y=box)
        it = dg2d_interp(**kv)
        x, y = F(n, n, m, m), F(n, n, m, m)
        it.call("get_coords", x, y, n, n, m, m)
        nodes = F(4, n, n, m, m)
        it.call("get_initial_conditions", x, y, nodes, n, n, m, m)
        tag = f"ninit{ninit}"
        out[f"{tag}/meta"] = np.array([ninit, n, m])
        out[f"{tag}/boxlen"] = np.array(box)
        out[f"{tag}/x"] = C(x); out[f"{tag}/y"] = C(y); out[f"{tag}/nodes"] = C(nodes)
        print(f"dg2d_ics {tag}: {time.time() - t0:.1f} s, |u|max {np.abs(nodes).max():.4g}", flush=True)
    np.savez_compressed(os.path.join(HERE, "ref_dg2d_ics.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["fv2d", "dg2d", "fv1d", "dg1d"]
    np.seterr(all="ignore")
    for w in which:
        if ":" in w:
            fam, tags = w.split(":")
            globals()[fam](only=tags.split(","))
        else:
            globals()[w]()
