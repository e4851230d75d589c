"""CPU tests of the 1D FV oracle (oracle/fv1d.c): fvm.f90 and benchmark_1d.f90 ('FVM' | 'EQL' | 'WB1')."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fv1d.npz")


def b1(o, solver, neq=2, ninit=2, nx=128, bc=2, eta=None):
    p = o.fv1d_params(nx=nx, solver=solver, nequilibrium=neq, bc=bc)
    x = o.fv1d_get_x(p)
    weq = o.fv1d_get_equilibrium_solution(p, x)
    u = o.fv1d_get_initial_conditions(p, ninit, x) if eta is None else o.fv1d_get_initial_conditions(p, ninit, x, eta)
    return p, x, weq, u


def test_eql_rhs_is_bitwise_zero_at_equilibrium(oracle):
    """1D FV-EQL gives an exactly zero RHS at the discrete equilibrium (benchmark_1d.f90:368-369)."""
    for neq, ninit in ((2, 1), (3, 3)):
        for nx in (32, 128, 513):
            p, x, weq, u = b1(oracle, "EQL", neq, ninit, nx)
            assert np.all(oracle.fv1d_compute_update(p, u, weq) == 0.0)
            un, it, t, dt = oracle.fv1d_evolve(p, u, weq, 0.05)
            assert it > 0 and np.array_equal(un, u)


def test_wb1_preserves_isentropic_state_to_roundoff_and_isothermal_to_second_order(oracle):
    """local hydrostatic reconstruction (benchmark_1d.f90:608-634): exact for the isentropic atmosphere;
    for the (default) isothermal one the face mismatch is O(dx^2), i.e. the RHS residual is O(dx)."""
    p, x, weq, u = b1(oracle, "WB1", 3, 3, 128)
    assert np.abs(oracle.fv1d_compute_update_sr(p, u, weq)).max() < 1e-11
    errs = []
    for nx in (64, 128, 256):
        p, x, weq, u = b1(oracle, "WB1", 2, 1, nx)
        errs.append(np.abs(oracle.fv1d_compute_update_sr(p, u, weq)[2:-2]).max())
    assert 0.9 < np.log2(errs[0] / errs[1]) < 1.1 and 0.9 < np.log2(errs[1] / errs[2]) < 1.1


def test_plain_fvm_is_not_balanced(oracle):
    p, x, weq, u = b1(oracle, "FVM", 2, 1, 128)
    assert np.abs(oracle.fv1d_compute_update_fvm(p, u, weq)).max() > 1e-3


def test_end_cells_copy_their_neighbours(oracle):
    for solver, fn in (("FVM", oracle.fv1d_compute_update_fvm), ("EQL", oracle.fv1d_compute_update), ("WB1", oracle.fv1d_compute_update_sr)):
        p, x, weq, u = b1(oracle, solver, 2, 2, 64, eta=1e-3)
        d = fn(p, u, weq)
        assert np.array_equal(d[0], d[1]) and np.array_equal(d[-1], d[-2])


def test_fvm_eql_second_stage_is_evaluated_at_u(oracle):
    """benchmark_1d.f90:228,:236 pass u (not w1) to the second RK stage; 'WB1' (:244) passes w1."""
    p, x, weq, u = b1(oracle, "EQL", 2, 2, 64, eta=1e-3)
    un, it, t, dt = oracle.fv1d_evolve(p, u, weq, 1.0, 1)
    d = oracle.fv1d_compute_update(p, u, weq)
    w1 = u + dt * d
    assert np.array_equal(un, 0.5 * u + 0.5 * w1 + 0.5 * dt * d)
    p, x, weq, u = b1(oracle, "WB1", 2, 2, 64, eta=1e-3)
    un, it, t, dt = oracle.fv1d_evolve(p, u, weq, 1.0, 1)
    d = oracle.fv1d_compute_update_sr(p, u, weq)
    w1 = u + dt * d
    assert np.array_equal(un, 0.5 * u + 0.5 * w1 + 0.5 * dt * oracle.fv1d_compute_update_sr(p, w1, weq))
    c = oracle.fv1d_compute_max_speed(p, u)
    assert dt == float(np.float32(0.8)) * (1.0 / 64) / c / 3.0


def test_fvm_f90_sod_and_periodic_conservation(oracle):
    p = oracle.fvm1d_params(nx=200, bc=2, source=1)
    u = oracle.fvm1d_initial_conditions(p, 4)
    un, it, t, dt = oracle.fvm1d_evolve(p, u, 0.1)
    assert t >= 0.1 and 0.12 < un[:, 0].min() and un[:, 0].max() <= 1.0 + 1e-12
    c = oracle.fvm1d_compute_max_speed(p, u)
    _, _, _, dt1 = oracle.fvm1d_evolve(p, u, 1.0, 1)
    assert dt1 == float(np.float32(0.8)) * (1.0 / 200) / c / 7.0          # n = 3 -> 2n+1 = 7 (fvm.f90:59)
    p = oracle.fvm1d_params(nx=100, bc=1, source=1)
    u = oracle.fvm1d_initial_conditions(p, 1)
    d = oracle.fvm1d_compute_update(p, u)
    assert np.abs(d.sum(axis=0)).max() < 1e-11                             # periodic + no source: telescoping fluxes
    # first-order convergence of the advected sine wave
    errs = []
    for nx in (100, 200):
        p = oracle.fvm1d_params(nx=nx, bc=1, source=1)
        u = oracle.fvm1d_initial_conditions(p, 1)
        un, it, t, dt = oracle.fvm1d_evolve(p, u, 0.1)
        xs = (np.arange(nx) + 0.5) / nx
        errs.append(np.abs(un[:, 0] - (1 + 0.5 * np.sin(2 * np.pi * (xs - t)))).max())
    assert 0.7 < np.log2(errs[0] / errs[1]) < 1.2


def test_golden_vectors(oracle):
    g = np.load(GOLD)
    for tag in [k[:-5] for k in g.files if k.endswith("_meta")]:
        kind = int(g[f"{tag}_meta"][0])
        u0 = g[f"{tag}_u0"]
        if kind == 0:
            _, nx, bc, source, steps = (int(v) for v in g[f"{tag}_meta"])
            p = oracle.fvm1d_params(nx=nx, bc=bc, source=source)
            assert np.array_equal(oracle.fvm1d_compute_update(p, u0), g[f"{tag}_dudt"])
            un, it, t, dt = oracle.fvm1d_evolve(p, u0, 1.0, steps)
        else:
            _, nx, bc, neq, solver, steps = (int(v) for v in g[f"{tag}_meta"])
            p = oracle.fv1d_params(nx=nx, bc=bc, nequilibrium=neq); p.solver = solver
            weq = g[f"{tag}_weq"]
            fn = {1: oracle.fv1d_compute_update_fvm, 2: oracle.fv1d_compute_update, 3: oracle.fv1d_compute_update_sr}[solver]
            assert np.array_equal(fn(p, u0, weq), g[f"{tag}_dudt"])
            un, it, t, dt = oracle.fv1d_evolve(p, u0, weq, 1.0, steps)
        assert np.array_equal(un, g[f"{tag}_un"]) and np.array_equal(np.array([it, t, dt]), g[f"{tag}_clock"])
