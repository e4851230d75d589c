"""CPU tests of the 2D FV oracle (oracle/fv2d.c): analytic invariants implied by benchmark_2d.f90,
an independent numpy restatement, and the committed golden vectors."""
import os

import numpy as np
import pytest

import np_fv2d

G32 = float(np.float32(1.4))
GOLD = os.path.join(os.path.dirname(__file__), "golden", "fv2d.npz")


def setup(o, nx, ny, ninit, neq=2, **kw):
    p = o.fv2d_params(nx, ny, neq, **kw)
    x, y = o.fv2d_get_coords(p)
    return p, o.fv2d_get_initial_conditions(p, ninit, x, y), o.fv2d_get_equilibrium_solution(p, x, y)


def fshape(a):  # (ny,nx,4) C-order == Fortran (4,nx,ny)
    return np.ascontiguousarray(a.transpose(2, 1, 0))


@pytest.mark.parametrize("nx,ny", [(16, 16), (32, 32), (64, 48), (50, 77)])
@pytest.mark.parametrize("neq,ninit", [(2, 2), (1, 1)])
def test_hydrostatic_state_gives_bitwise_zero_rhs(oracle, nx, ny, neq, ninit):
    """IC == equilibrium -> dudt == 0 bit for bit (benchmark_2d.f90:599-609).
    Holds whenever the discrete y-balance is within a factor 2 of the source (Sterbenz), i.e.
    0.5 < dx/dy < 2 given the reference's y_faces=(j-1)*dx (:513)."""
    p, u, weq = setup(oracle, nx, ny, ninit, neq)
    d = oracle.fv2d_compute_update_exact(p, u, weq)
    assert np.all(d == 0.0)
    un, it, t, dt, cm = oracle.fv2d_evolve(p, u, weq, 1.0, 4)
    assert it == 4 and np.array_equal(un, u)


def test_hydrostatic_state_strongly_anisotropic_grid_is_roundoff_only(oracle):
    """dx/dy > 2: y_faces=(j-1)*dx (:513) puts the 'face' equilibrium far from the face, the Sterbenz
    cancellation no longer applies and the reference's RHS is round-off instead of exactly 0."""
    p, u, weq = setup(oracle, 37, 91, 2, 2)
    d = oracle.fv2d_compute_update_exact(p, u, weq)
    assert 0 < np.abs(d).max() < 1e-12


@pytest.mark.parametrize("nx,ny,ninit,neq", [(24, 24, 3, 2), (33, 20, 3, 2), (16, 40, 4, 2), (20, 24, 1, 1), (3, 3, 3, 2)])
def test_c_oracle_equals_numpy_restatement_bitwise(oracle, nx, ny, ninit, neq):
    p, u, weq = setup(oracle, nx, ny, ninit, neq)
    d = oracle.fv2d_compute_update_exact(p, u, weq)
    dn = np_fv2d.compute_update_exact(fshape(u), fshape(weq), nx, ny, neq, np.float64(G32)).transpose(2, 1, 0)
    assert np.array_equal(d, dn)


def test_boundary_lines_are_frozen(oracle):
    p, u, weq = setup(oracle, 20, 28, 4)
    d = oracle.fv2d_compute_update_exact(p, u, weq)
    assert np.all(d[0] == 0) and np.all(d[-1] == 0) and np.all(d[:, 0] == 0) and np.all(d[:, -1] == 0)
    assert np.abs(d[1:-1, 1:-1]).max() > 0


def test_literal_kinds(oracle):
    """gamma, 1.21, eta are real(4) literals promoted to real(8); (i-0.5) is single precision."""
    p, u, weq = setup(oracle, 8, 8, 2)
    assert p.gamma == 1.399999976158142
    x, y = oracle.fv2d_get_coords(p)
    assert x[0, 0] == 0.5 * (1.0 / 8) and y[3, 0] == 3.5 / 8
    rho0 = float(np.float32(1.21))
    import math
    assert weq[0, 0, 0] == rho0 * math.exp(-(rho0) * (x[0, 0] + y[0, 0]))
    assert weq[0, 0, 3] == math.exp(-(rho0) * (x[0, 0] + y[0, 0]))


def test_max_speed_and_dt(oracle):
    p, u, weq = setup(oracle, 40, 30, 3)
    c = oracle.fv2d_compute_max_speed(p, u)
    cn = np_fv2d.compute_speed(fshape(u), np.float64(G32)).max()
    assert c == cn
    un, it, t, dt, cm = oracle.fv2d_evolve(p, u, weq, 1.0, 1)
    assert cm == c and dt == 0.5 * (1.0 / 40) / c * 0.5 and t == dt


def test_evolve_is_ssp_rk2_of_the_update(oracle):
    """benchmark_2d.f90:246-250 restated with numpy on top of the oracle's RHS."""
    p, u, weq = setup(oracle, 28, 28, 3)
    un, it, t, dt, cm = oracle.fv2d_evolve(p, u, weq, 1.0, 1)
    d1 = oracle.fv2d_compute_update_exact(p, u, weq)
    w1 = u + dt * d1
    d2 = oracle.fv2d_compute_update_exact(p, w1, weq)
    assert np.array_equal(un, 0.5 * u + 0.5 * w1 + 0.5 * dt * d2)


def test_evolve_overshoots_tend_like_the_reference(oracle):
    p, u, weq = setup(oracle, 16, 16, 3)
    un, it, t, dt, cm = oracle.fv2d_evolve(p, u, weq, 0.02, -1)
    assert t >= 0.02 and t - dt < 0.02 and it >= 2


def test_perturbation_scales_linearly(oracle):
    """Well-balancedness in practice: the RHS is O(eta) (no O(dx) equilibrium truncation error)."""
    p = oracle.fv2d_params(48, 48)
    x, y = oracle.fv2d_get_coords(p)
    weq = oracle.fv2d_get_equilibrium_solution(p, x, y)
    d = [np.abs(oracle.fv2d_compute_update_exact(p, oracle.fv2d_get_initial_conditions(p, 3, x, y, eta=e), weq)).max()
         for e in (1e-5, 1e-8)]
    assert d[0] / d[1] == pytest.approx(1e3, rel=1e-3)


def test_plain_scheme_is_not_well_balanced(oracle):
    p, u, weq = setup(oracle, 32, 32, 2)
    assert np.abs(oracle.fv2d_compute_update(p, u, weq)).max() > 1e-6


def test_threads_do_not_change_results(oracle):
    p, u, weq = setup(oracle, 40, 36, 3)
    oracle.set_num_threads(1)
    a = oracle.fv2d_evolve(p, u, weq, 1.0, 2)[0]
    oracle.set_num_threads(4)
    b = oracle.fv2d_evolve(p, u, weq, 1.0, 2)[0]
    assert np.array_equal(a, b)


def test_golden_vectors(oracle):
    g = np.load(GOLD)
    for tag in ("sq_pert", "ragged_pert", "riemann", "riemann_ragged", "eq1"):
        nx, ny, ninit, neq = (int(v) for v in g[f"{tag}_meta"])
        p = oracle.fv2d_params(nx, ny, neq)
        u, weq = g[f"{tag}_u"], g[f"{tag}_weq"]
        assert np.array_equal(oracle.fv2d_compute_update_exact(p, u, weq), g[f"{tag}_dudt"])
        assert np.array_equal(oracle.fv2d_compute_update(p, u, weq), g[f"{tag}_dudt_plain"])
        un, it, t, dt, cm = oracle.fv2d_evolve(p, u, weq, 1.0, 3)
        assert np.array_equal(un, g[f"{tag}_u3"])
        assert np.array_equal(np.array([it, t, dt, cm]), g[f"{tag}_clock"])
