"""Generates the committed golden vectors under tests/golden/ from the C oracle (oracle/*.c).

The reference itself cannot be run here (no Fortran compiler in the image), so these vectors pin the
ORACLE against regressions; the oracle in turn is pinned by the analytic invariants in
tests/test_oracle_*.py and by independent numpy restatements.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wb_oracle as o  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fv2d():
    out = {}
    for tag, nx, ny, ninit, neq in (("sq_pert", 24, 24, 3, 2), ("ragged_pert", 33, 20, 3, 2),
                                    ("riemann", 40, 40, 4, 2), ("riemann_ragged", 32, 24, 4, 2), ("eq1", 20, 24, 1, 1)):
        p = o.fv2d_params(nx, ny, neq)
        x, y = o.fv2d_get_coords(p)
        weq = o.fv2d_get_equilibrium_solution(p, x, y)
        u = o.fv2d_get_initial_conditions(p, ninit, x, y)
        out[f"{tag}_meta"] = np.array([nx, ny, ninit, neq])
        out[f"{tag}_u"] = u
        out[f"{tag}_weq"] = weq
        out[f"{tag}_dudt"] = o.fv2d_compute_update_exact(p, u, weq)
        out[f"{tag}_dudt_plain"] = o.fv2d_compute_update(p, u, weq)
        un, it, t, dt, cm = o.fv2d_evolve(p, u, weq, 1.0, 3)
        out[f"{tag}_u3"] = un
        out[f"{tag}_clock"] = np.array([it, t, dt, cm])
    np.savez_compressed(os.path.join(HERE, "fv2d.npz"), **out)


def dg2d():
    """2D DG: shipped pulse (ninit=1), hydrostatic+bump with gravity source (ninit=2, source=2, grad_phi_case=1 and 2),
    Riemann problem with each limiter; orders 1-3; RK4 / EQL / DEB."""
    out = {}
    cases = [  # tag, nx, mx, bc, source, gcase, flux, limiter, solver, ninit, steps
        ("pulse_o2_onp", 8, 2, 1, 1, 2, "llf1", "ONP", "RK4", 1, 3),
        ("pulse_o3_onp", 6, 3, 1, 1, 2, "llf1", "ONP", "RK4", 1, 2),
        ("pulse_o1", 8, 1, 1, 1, 2, "llf1", "ONP", "RK4", 1, 3),
        ("shipped_flux", 6, 2, 1, 1, 2, "llf", "ONP", "RK4", 1, 2),
        ("hydro_o3_g1", 6, 3, 2, 2, 1, "llf1", "ONP", "RK4", 2, 2),
        ("hydro_o2_kepler", 6, 2, 2, 2, 2, "llf1", "ONP", "EQL", 2, 2),
        ("riemann_o3_hio", 8, 3, 2, 1, 2, "llf1", "HIO", "RK4", 3, 3),
        ("riemann_o2_1or", 8, 2, 2, 1, 2, "llf1", "1OR", "EQL", 4, 3),
        ("riemann_o3_low", 6, 3, 2, 1, 2, "llf1", "LOW", "DEB", 4, 3),
        ("advsink_o2", 6, 2, 1, 3, 2, "llf1", "none", "SS4", 1, 2),
        ("riemann_o4_onp", 4, 4, 3, 1, 2, "llf1", "ONP", "RK4", 5, 2),
        ("pulse_o3_hll2", 6, 3, 1, 1, 2, "hll2", "ONP", "RK4", 1, 2),
        ("riemann_o2_hllc", 8, 2, 2, 1, 2, "hllc", "ONP", "RK4", 3, 2),
    ]
    for tag, nx, mx, bc, source, gcase, flux, lim, solver, ninit, steps in cases:
        p = o.dg2d_params(nx=nx, ny=nx, mx=mx, my=mx, bc=bc, source=source, grad_phi_case=gcase, flux=flux, limiter=lim,
                          solver=solver, ninit=ninit)
        x, y = o.dg2d_get_coords(p)
        u0 = o.dg2d_get_initial_conditions(p, x, y)
        m0 = o.dg2d_get_modes_from_nodes(p, u0)
        out[f"{tag}_meta"] = np.array([nx, mx, bc, source, gcase, p.flux_id, p.limiter_id, p.solver_id, ninit, steps])
        out[f"{tag}_u0"] = u0
        out[f"{tag}_dudt"] = o.dg2d_compute_update(p, m0, x, y)
        out[f"{tag}_lim"] = o.dg2d_apply_limiter(p, m0)
        un, it, t, dt = o.dg2d_evolve(p, u0, x, y, 1.0, steps)
        out[f"{tag}_un"] = un
        out[f"{tag}_clock"] = np.array([it, t, dt])
        assert np.all(np.isfinite(un)), tag
    np.savez_compressed(os.path.join(HERE, "dg2d.npz"), **out)


def fv1d():
    out = {}
    for tag, nx, bc, source, ninit, steps in (("fvm_sod", 200, 2, 2, 4, 5), ("fvm_sine_periodic", 64, 1, 1, 1, 5),
                                              ("fvm_hydro", 50, 2, 2, 7, 4)):
        p = o.fvm1d_params(nx=nx, bc=bc, source=source)
        u0 = o.fvm1d_initial_conditions(p, ninit)
        out[f"{tag}_meta"] = np.array([0, nx, bc, source, steps])
        out[f"{tag}_u0"] = u0
        out[f"{tag}_dudt"] = o.fvm1d_compute_update(p, u0)
        un, it, t, dt = o.fvm1d_evolve(p, u0, 1.0, steps)
        out[f"{tag}_un"] = un; out[f"{tag}_clock"] = np.array([it, t, dt])
    for tag, nx, bc, neq, solver, ninit, eta, steps in (("b1_wb1_default", 128, 2, 2, "WB1", 2, o.F32(1e-8), 5),
                                                        ("b1_wb1_isentropic", 64, 2, 3, "WB1", 3, 0.0, 4),
                                                        ("b1_eql_bump", 96, 2, 2, "EQL", 2, 1e-3, 5),
                                                        ("b1_eql_bc3", 40, 3, 2, "EQL", 2, 1e-3, 3),
                                                        ("b1_fvm_bump", 77, 1, 2, "FVM", 2, 1e-3, 4)):
        p = o.fv1d_params(nx=nx, bc=bc, nequilibrium=neq, solver=solver)
        x = o.fv1d_get_x(p); weq = o.fv1d_get_equilibrium_solution(p, x); u0 = o.fv1d_get_initial_conditions(p, ninit, x, eta)
        fn = {"FVM": o.fv1d_compute_update_fvm, "EQL": o.fv1d_compute_update, "WB1": o.fv1d_compute_update_sr}[solver]
        out[f"{tag}_meta"] = np.array([1, nx, bc, neq, p.solver, steps])
        out[f"{tag}_u0"] = u0; out[f"{tag}_weq"] = weq
        out[f"{tag}_dudt"] = fn(p, u0, weq)
        un, it, t, dt = o.fv1d_evolve(p, u0, weq, 1.0, steps)
        out[f"{tag}_un"] = un; out[f"{tag}_clock"] = np.array([it, t, dt])
    np.savez_compressed(os.path.join(HERE, "fv1d.npz"), **out)


def dg1d():
    out = {}
    for tag, n, nx, riemann, source, ninit, pert, steps in (("default_hllc", 3, 128, 2, 2, 8, o.F32(1e-8), 4),
                                                           ("llf_o2", 2, 64, 1, 2, 8, 1e-3, 4),
                                                           ("hllc_o1", 1, 40, 2, 2, 8, 1e-2, 5),
                                                           ("sod_nosource", 3, 50, 2, 1, 4, 0.0, 3),
                                                           ("steady_hllc", 3, 32, 2, 2, 7, 0.0, 3)):
        p = o.dg1d_params(n=n, nx=nx, riemann=riemann, source=source, ninit=ninit, pert=pert)
        ui, ueq, du = o.dg1d_setup(p)
        out[f"{tag}_meta"] = np.array([n, nx, riemann, source, steps])
        out[f"{tag}_du"] = du; out[f"{tag}_ueq"] = ueq; out[f"{tag}_ui"] = ui
        out[f"{tag}_dudt"] = o.dg1d_compute_update_exact_delta(p, du, ueq)
        du2, ui2, it, t, dt = o.dg1d_evolve_rki(p, du, ueq, ui, 1.0, steps)
        out[f"{tag}_du2"] = du2; out[f"{tag}_ui2"] = ui2; out[f"{tag}_clock"] = np.array([it, t, dt])
        assert np.all(np.isfinite(du2)), tag
    # plain update / limiter / 'RK1'..'RK4' (dg_with_source.f90:807-1028, :414-519, :173-230)
    rng = np.random.default_rng(11)
    for tag, n, nx, riemann, source, bc, use_limiter, integ, steps, ninit, pert in (
            ("rk2_periodic_lim", 3, 48, 2, 2, 1, 1, 2, 3, 8, 1e-2), ("rk3_zerograd", 3, 40, 1, 2, 2, 1, 3, 3, 8, 1e-3),
            ("rk4_reflect", 2, 33, 2, 2, 3, 1, 4, 2, 8, 1e-3), ("rk1_sod_bc2", 3, 50, 2, 1, 2, 0, 1, 4, 4, 0.0),
            ("rk3_sod_lim", 3, 50, 2, 1, 2, 1, 3, 4, 4, 0.0), ("rk2_bc4", 2, 20, 1, 2, 4, 0, 2, 2, 8, 1e-2)):
        p = o.dg1d_params(n=n, nx=nx, riemann=riemann, source=source, ninit=ninit, pert=pert, bc=bc, use_limiter=use_limiter)
        ui, ueq, du = o.dg1d_setup(p)
        u = o.dg1d_project(p, ui)
        if n > 1:
            u[:, 1:, :] += 1e-3 * rng.standard_normal(u[:, 1:, :].shape)      # exercise the moment limiter
        out[f"{tag}_pmeta"] = np.array([n, nx, riemann, source, bc, use_limiter, integ, steps])
        out[f"{tag}_u"] = u; out[f"{tag}_du"] = du; out[f"{tag}_ueq"] = ueq; out[f"{tag}_ui"] = ui
        out[f"{tag}_dudt"] = o.dg1d_compute_update(p, u)
        out[f"{tag}_lim"] = o.dg1d_limiter(p, u)
        u2, ui2, it, t, dt = o.dg1d_evolve_rk(p, f"RK{integ}", u, du, ueq, ui, 1.0, steps)
        out[f"{tag}_u2"] = u2; out[f"{tag}_ui2"] = ui2; out[f"{tag}_pclock"] = np.array([it, t, dt])
        assert np.all(np.isfinite(u2)), tag
    # 'RKw' / 'RKe': compute_update_exact, limiter_TDV, limiter_cons (dg_with_source.f90:1380-1744, :520-734, :229-280)
    for tag, n, nx, riemann, source, bc, use_limiter, integ, steps, ninit, pert in (
            ("rkw_default_bc5", 3, 64, 2, 2, 5, 0, 5, 3, 8, 1e-3), ("rkw_llf_bc4", 2, 40, 1, 2, 4, 0, 5, 3, 8, 1e-2),
            ("rke_bc5", 3, 64, 2, 2, 5, 0, 6, 4, 8, 1e-3), ("rke_lim_bc2", 3, 48, 1, 2, 2, 1, 6, 3, 8, 1e-2),
            ("rke_o1", 1, 32, 2, 2, 5, 0, 6, 3, 8, 1e-2)):
        p = o.dg1d_params(n=n, nx=nx, riemann=riemann, source=source, ninit=ninit, pert=pert, bc=bc, use_limiter=use_limiter)
        ui, ueq, du = o.dg1d_setup(p)
        u = o.dg1d_project(p, ui); q = o.dg1d_project(p, ueq)
        lin = u.copy()
        if n > 1:
            lin[:, 1:, :] += 1e-3 * rng.standard_normal(lin[:, 1:, :].shape)
            lin[3, 1, 0] = 5.0                                                 # a cell with a negative density trace
        out[f"{tag}_wmeta"] = np.array([n, nx, riemann, source, bc, use_limiter, integ, steps])
        out[f"{tag}_u"] = u; out[f"{tag}_du"] = du; out[f"{tag}_ueq"] = ueq; out[f"{tag}_q"] = q; out[f"{tag}_ui"] = ui
        out[f"{tag}_lin"] = lin
        if bc in (4, 5):
            out[f"{tag}_dudt"] = o.dg1d_compute_update_exact(p, u, q)
        out[f"{tag}_lcons"] = o.dg1d_limiter_cons(p, lin)
        if not use_limiter:
            out[f"{tag}_ltdv"] = o.dg1d_limiter_tdv(p, lin)
        uu, dd, ui2, it, t, dt = o.dg1d_evolve_w(p, "RKw" if integ == 5 else "RKe", u, du, ueq, q, ui, 1.0, steps)
        out[f"{tag}_u2"] = uu; out[f"{tag}_du2"] = dd; out[f"{tag}_ui2"] = ui2; out[f"{tag}_wclock"] = np.array([it, t, dt])
        assert np.all(np.isfinite(uu)) and np.all(np.isfinite(dd)), tag
    np.savez_compressed(os.path.join(HERE, "dg1d.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["fv2d", "dg2d", "fv1d", "dg1d"]
    for w in which:
        globals()[w]()
        print("wrote", w)
