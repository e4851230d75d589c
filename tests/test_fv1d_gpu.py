"""GPU parity tests of the 1D FV paths (fvm.f90, benchmark_1d.f90) against the CPU oracle.

The kernels keep the reference's operation order; the only differences to the CPU restatement are exp()/pow()
(CUDA vs libm, <= 2 ulp), so: bit-for-bit where no transcendental enters the RHS (fvm.f90, 'FVM'), <= 1e-12 relative
L-inf otherwise, and the well-balanced invariants exactly as on the CPU."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "fv1d.npz")
TOL = 1e-12


@pytest.fixture(scope="module")
def wb():
    import __graft_entry__ as ge
    ge.build()
    import wbeuler
    return wbeuler


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("nx,bc,source,ninit", [(200, 2, 2, 4), (64, 1, 1, 1), (50, 2, 2, 7), (3, 2, 2, 4), (513, 1, 2, 3)])
def test_fvm_f90_bitwise(wb, oracle, nx, bc, source, ninit):
    p = oracle.fvm1d_params(nx=nx, bc=bc, source=source)
    u = oracle.fvm1d_initial_conditions(p, ninit)
    with wb.FVM1D(nx=nx, bc=bc, source=source) as s:
        assert np.array_equal(s.compute_update(u), oracle.fvm1d_compute_update(p, u))
        assert s.compute_max_speed(u) == oracle.fvm1d_compute_max_speed(p, u)
        got, it, t, dt = s.evolve(u, 1.0, 7)
    ref, it0, t0, dt0 = oracle.fvm1d_evolve(p, u, 1.0, 7)
    assert (it, t, dt) == (it0, t0, dt0) and np.array_equal(got, ref)


def test_fvm_f90_until_tend(wb, oracle):
    p = oracle.fvm1d_params()
    u = oracle.fvm1d_initial_conditions(p, 4)
    ref, it0, t0, dt0 = oracle.fvm1d_evolve(p, u, 0.02)
    with wb.FVM1D() as s:
        got, it, t, dt = s.evolve(u, 0.02)
    assert (it, t) == (it0, t0) and t >= 0.02 and np.array_equal(got, ref)


def setup_b1(o, solver, neq, ninit, nx, bc=2, eta=1e-3):
    p = o.fv1d_params(nx=nx, solver=solver, nequilibrium=neq, bc=bc)
    x = o.fv1d_get_x(p)
    return p, o.fv1d_get_equilibrium_solution(p, x), o.fv1d_get_initial_conditions(p, ninit, x, eta)


@pytest.mark.parametrize("solver", ["FVM", "EQL", "WB1"])
@pytest.mark.parametrize("nx,bc,neq,ninit", [(128, 2, 2, 2), (64, 1, 2, 2), (40, 3, 2, 2), (96, 2, 3, 3), (4, 2, 2, 2)])
def test_benchmark_1d_rhs_and_evolve(wb, oracle, solver, nx, bc, neq, ninit):
    p, weq, u = setup_b1(oracle, solver, neq, ninit, nx, bc)
    ofn = {"FVM": oracle.fv1d_compute_update_fvm, "EQL": oracle.fv1d_compute_update, "WB1": oracle.fv1d_compute_update_sr}[solver]
    with wb.FV1D(nx=nx, bc=bc, nequilibrium=neq, solver=solver) as s:
        gfn = {"FVM": s.compute_update_fvm, "EQL": s.compute_update, "WB1": s.compute_update_sr}[solver]
        d, dref = gfn(u, weq), ofn(p, u, weq)
        c = oracle.fv1d_compute_max_speed(p, u)
        assert s.compute_max_speed(u) == c
        dt = float(np.float32(0.8)) * (1.0 / nx) / c / 3.0
        assert np.abs(dt * (d - dref)).max() / np.abs(u).max() <= TOL
        if solver == "FVM":
            assert np.array_equal(d, dref)          # no exp/pow in the plain scheme
        got, it, t, dtl = s.evolve(u, weq, 1.0, 6)
    ref, it0, t0, dt0 = oracle.fv1d_evolve(p, u, weq, 1.0, 6)
    assert it == it0 == 6 and abs(t - t0) <= 1e-14 * t0
    assert rel(got, ref) <= TOL


def test_eql_is_exactly_well_balanced_on_the_gpu(wb, oracle):
    for neq, ninit in ((2, 1), (3, 3)):
        p, weq, u = setup_b1(oracle, "EQL", neq, ninit, 128)
        with wb.FV1D(nx=128, nequilibrium=neq, solver="EQL") as s:
            assert np.all(s.compute_update(u, weq) == 0.0)
            got, it, t, dt = s.evolve(u, weq, 0.05)
            assert it > 0 and np.array_equal(got, u)


def test_wb1_preserves_the_isentropic_atmosphere_to_roundoff(wb, oracle):
    p, weq, u = setup_b1(oracle, "WB1", 3, 3, 128)
    with wb.FV1D(nx=128, nequilibrium=3, solver="WB1") as s:
        assert np.abs(s.compute_update_sr(u)).max() < 1e-11
        got, it, t, dt = s.evolve(u, weq, 0.2)
    assert t >= 0.2 and np.abs(got - u).max() < 1e-13


def test_default_configuration_until_tend(wb, oracle):
    """parameters.f90 as shipped: nx=128, WB1, ninit=2, eta=1e-8 (real(4)), tend=0.2."""
    p = oracle.fv1d_params()
    x = oracle.fv1d_get_x(p); weq = oracle.fv1d_get_equilibrium_solution(p, x); u = oracle.fv1d_get_initial_conditions(p, 2, x)
    ref, it0, t0, dt0 = oracle.fv1d_evolve(p, u, weq, 0.2)
    with wb.FV1D() as s:
        got, it, t, dt = s.evolve(u, weq, 0.2)
    assert it == it0 and abs(t - t0) <= 1e-13 * t0 and rel(got, ref) <= TOL


def test_golden_vectors(wb):
    g = np.load(GOLD)
    inv = {1: "FVM", 2: "EQL", 3: "WB1"}
    for tag in [k[:-5] for k in g.files if k.endswith("_meta")]:
        kind = int(g[f"{tag}_meta"][0]); u0 = g[f"{tag}_u0"]
        if kind == 0:
            _, nx, bc, source, steps = (int(v) for v in g[f"{tag}_meta"])
            with wb.FVM1D(nx=nx, bc=bc, source=source) as s:
                got, it, t, dt = s.evolve(u0, 1.0, steps)
            assert np.array_equal(got, g[f"{tag}_un"]), tag
        else:
            _, nx, bc, neq, solver, steps = (int(v) for v in g[f"{tag}_meta"])
            with wb.FV1D(nx=nx, bc=bc, nequilibrium=neq, solver=inv[solver]) as s:
                got, it, t, dt = s.evolve(u0, g[f"{tag}_weq"], 1.0, steps)
            assert rel(got, g[f"{tag}_un"]) <= TOL, tag
        assert it == int(g[f"{tag}_clock"][0])
