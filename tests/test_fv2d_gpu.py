"""GPU parity tests of the 2D well-balanced FV path (C-ABI -> CUDA) against the CPU oracle.

Tolerances (BASELINE.json north_star): conserved fields within 1e-12 relative L-inf of the reference
arithmetic; hydrostatic state preserved exactly (bitwise-zero RHS) wherever the reference does.
For a bare RHS the 1e-12 bound is applied to the field it produces: |dt*(dudt - dudt_ref)| / max|u| with the
CFL dt of the state, because dudt itself is a difference of O(1) fluxes divided by dx."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fv2d.npz")
TOL = 1e-12


@pytest.fixture(scope="module")
def wb():
    import __graft_entry__ as ge
    ge.build()
    import wbeuler
    return wbeuler


def setup(o, nx, ny, ninit, neq=2, **kw):
    p = o.fv2d_params(nx, ny, neq, **kw)
    x, y = o.fv2d_get_coords(p)
    return p, o.fv2d_get_initial_conditions(p, ninit, x, y), o.fv2d_get_equilibrium_solution(p, x, y)


def rel_linf_fields(a, b):
    """Parity metric.  For every conserved field v: Linf(a_v - b_v) / Linf(b) with Linf(b) the norm of the whole
    conserved state -- a field that is (nearly) zero, like the momenta of a slightly perturbed atmosphere
    (|m| ~ 1e-6), carries the reference's own O(ulp(p)/dx*dt) rounding noise, which no re-ordered evaluation
    can reproduce digit for digit (the reference-order kernel differs from the oracle there too, through the
    last bit of exp() alone).  Fields whose own norm is within 1e-3 of the state norm (always rho and E; the
    momenta in the Riemann problem) are also held to 1e-12 of their OWN norm.  Fields that are identically
    zero must match exactly."""
    scale = np.abs(b).max()
    worst = 0.0
    for v in range(4):
        den = np.abs(b[..., v]).max()
        num = np.abs(a[..., v] - b[..., v]).max()
        if den == 0.0:
            assert num == 0.0
            continue
        worst = max(worst, num / scale)
        if den >= 1e-3 * scale:
            worst = max(worst, num / den)
    return worst


def own_norm_errs(a, b):
    """Linf(a_v - b_v) / Linf(b_v) for every conserved field with a non-zero norm: the error of a field against ITS OWN
    size (the momenta of the 1e-5 pressure bump are ~1e-7 of the state and are the only signal of the bump)."""
    out = []
    for v in range(4):
        den = np.abs(b[..., v]).max()
        out.append(np.abs(a[..., v] - b[..., v]).max() / den if den > 0 else 0.0)
    return out


# Stated bound for the own-norm error of the small fields: the reference's own rounding noise in the momenta is
# O(ulp(p)/dx * dt) per stage ~ 1e-16 absolute, i.e. ~1e-9 of momenta of size 1e-7; a regression of the fused kernels from
# there to 1e-7 would still pass the state-norm bar, so it is asserted separately.  Measured on B200 (256^2, 4 steps):
# 2.8e-10 for the fused kernels, 1.8e-10 for the reference-order kernels (CUDA vs glibc exp alone).
OWN_NORM_TOL = 1e-9


def rhs_err(o, p, u, d, dref):
    dt = 0.5 * (p.boxlen_x / p.nx) / o.fv2d_compute_max_speed(p, u) * p.cfl
    return np.abs(dt * (d - dref)).max() / np.abs(u).max()


SIZES = [(3, 3), (8, 5), (24, 24), (33, 20), (64, 64), (100, 37), (257, 130)]


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("nx,ny", SIZES)
def test_compute_update_exact_matches_oracle(wb, oracle, nx, ny, arith):
    p, u, weq = setup(oracle, nx, ny, 3)
    dref = oracle.fv2d_compute_update_exact(p, u, weq)
    with wb.FV2D(nx, ny, arith=arith) as s:
        d = s.compute_update_exact(u, weq)
    assert np.all(np.isfinite(d))
    assert rhs_err(oracle, p, u, d, dref) <= TOL
    # boundary lines frozen exactly
    assert np.all(d[0] == 0) and np.all(d[-1] == 0) and np.all(d[:, 0] == 0) and np.all(d[:, -1] == 0)


@pytest.mark.parametrize("nx,ny", [(24, 24), (100, 37)])
def test_reference_order_kernel_is_nearly_bitwise(wb, oracle, nx, ny):
    """arith=1 runs the reference's operation order; the only difference left is libm vs CUDA exp (<=1 ulp)
    in the face equilibria, which the well-balanced form is insensitive to."""
    p, u, weq = setup(oracle, nx, ny, 3)
    dref = oracle.fv2d_compute_update_exact(p, u, weq)
    with wb.FV2D(nx, ny, arith=1) as s:
        d = s.compute_update_exact(u, weq)
    assert rhs_err(oracle, p, u, d, dref) <= 1e-15


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("nx,ny,neq,ninit", [(32, 32, 2, 2), (64, 48, 2, 2), (130, 130, 1, 1), (50, 77, 2, 2)])
def test_hydrostatic_state_rhs_is_bitwise_zero(wb, oracle, nx, ny, neq, ninit, arith):
    p, u, weq = setup(oracle, nx, ny, ninit, neq)
    with wb.FV2D(nx, ny, nequilibrium=neq, arith=arith) as s:
        d = s.compute_update_exact(u, weq)
        assert np.all(d == 0.0)
        un, it, t, dt = s.evolve(u, weq, 1.0, 10)
        assert it == 10 and np.array_equal(un, u)


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("nx,ny,ninit,steps", [(64, 64, 3, 10), (96, 130, 3, 6), (40, 40, 4, 8), (48, 36, 4, 6), (256, 256, 3, 4)])
def test_evolve_matches_oracle(wb, oracle, nx, ny, ninit, steps, arith):
    p, u, weq = setup(oracle, nx, ny, ninit)
    ref, it, t, dt, cm = oracle.fv2d_evolve(p, u, weq, 1.0, steps)
    with wb.FV2D(nx, ny, arith=arith) as s:
        got, it2, t2, dt2 = s.evolve(u, weq, 1.0, steps)
    assert it2 == it == steps
    assert abs(t2 - t) <= 1e-13 * t and abs(dt2 - dt) <= 1e-13 * dt
    assert rel_linf_fields(got, ref) <= TOL
    own = own_norm_errs(got, ref)
    print(f"own-norm errors (rho, mx, my, E) {nx}x{ny} ninit={ninit} arith={arith}: " + " ".join(f"{e:.2e}" for e in own))
    assert max(own) <= OWN_NORM_TOL


def test_evolve_until_tend_overshoots_like_the_reference(wb, oracle):
    p, u, weq = setup(oracle, 32, 32, 3)
    ref, it, t, dt, cm = oracle.fv2d_evolve(p, u, weq, 0.05, -1)
    with wb.FV2D(32, 32) as s:
        got, it2, t2, dt2 = s.evolve(u, weq, 0.05, -1)
    assert it2 == it and t2 >= 0.05 and abs(t2 - t) <= 1e-13 * t
    assert rel_linf_fields(got, ref) <= TOL


@pytest.mark.parametrize("arith", [0, 1])
def test_compute_max_speed(wb, oracle, arith):
    p, u, weq = setup(oracle, 70, 45, 4)
    with wb.FV2D(70, 45, arith=arith) as s:
        c = s.compute_max_speed(u)
    cref = oracle.fv2d_compute_max_speed(p, u)
    assert abs(c - cref) <= (4e-16 if arith == 0 else 0.0) * cref


def test_plain_compute_update(wb, oracle):
    p, u, weq = setup(oracle, 48, 40, 4)
    dref = oracle.fv2d_compute_update(p, u, weq)
    with wb.FV2D(48, 40, arith=1) as s:
        d = s.compute_update(u, weq)
    assert rhs_err(oracle, p, u, d, dref) <= 1e-15


def test_golden_vectors(wb):
    g = np.load(GOLD)
    for tag in ("sq_pert", "ragged_pert", "riemann", "riemann_ragged", "eq1"):
        nx, ny, ninit, neq = (int(v) for v in g[f"{tag}_meta"])
        u, weq = g[f"{tag}_u"], g[f"{tag}_weq"]
        with wb.FV2D(nx, ny, nequilibrium=neq) as s:
            got, it, t, dt = s.evolve(u, weq, 1.0, 3)
        assert rel_linf_fields(got, g[f"{tag}_u3"]) <= TOL
        assert it == int(g[f"{tag}_clock"][0]) and abs(t - g[f"{tag}_clock"][1]) <= 1e-13 * t


def test_device_initial_conditions_match_reference_formulae(wb, oracle):
    for ninit in (1, 2, 3, 4):
        p, u, weq = setup(oracle, 64, 40, ninit)
        with wb.FV2D(64, 40) as s:
            ud, wd = s.get_initial_conditions(ninit)
        assert np.abs(ud - u).max() <= 4e-16 * np.abs(u).max()     # libm vs CUDA exp
        assert np.abs(wd - weq).max() <= 4e-16 * np.abs(weq).max()


def test_resident_path_equals_evolve(wb, oracle):
    p, u, weq = setup(oracle, 128, 96, 3)
    with wb.FV2D(128, 96) as s:
        a, it, t, dt = s.evolve(u, weq, 1.0, 7)
        s.upload(u, weq)
        s.step_async(3); s.step_async(4)
        it2, t2, dt2, cm2 = s.sync()
        b = s.download()
    assert it2 == 7 and t2 == t and np.array_equal(a, b)


@pytest.mark.parametrize("n", [1024])
def test_large_grid_rhs_against_oracle(wb, oracle, n):
    p, u, weq = setup(oracle, n, n, 3)
    dref = oracle.fv2d_compute_update_exact(p, u, weq)
    with wb.FV2D(n, n) as s:
        d = s.compute_update_exact(u, weq)
    assert rhs_err(oracle, p, u, d, dref) <= TOL


def test_full_size_4096_properties(wb):
    """BASELINE config 3 size: properties that need no CPU run.
    (1) device-initialised hydrostatic state stays bitwise unchanged; (2) fused and reference-order kernels
    agree to 1e-12 on the perturbed state; (3) the x<->y mirror symmetry of the problem is kept."""
    n = 4096
    with wb.FV2D(n, n) as s:
        s.init_device(2)
        u0 = s.download()
        s.step_async(3)
        it, t, dt, cm = s.sync()
        assert it == 3 and np.array_equal(s.download(), u0)
        del u0
        s.init_device(3)
        s.step_async(3)
        s.sync()
        fast = s.download()
    with wb.FV2D(n, n, arith=1) as s:
        s.init_device(3)
        s.step_async(3)
        s.sync()
        ref = s.download()
    assert rel_linf_fields(fast, ref) <= TOL
    sym = fast.transpose(1, 0, 2)[..., [0, 2, 1, 3]]
    assert rel_linf_fields(sym, fast) <= TOL


@pytest.mark.parametrize("nx,ny", [(64, 48), (100, 37)])
def test_floors_of_the_sound_speed_are_honoured(wb, oracle, nx, ny):
    """compute_speed (benchmark_2d.f90:283-295) floors p and rho at 1d-10.  The fused kernels evaluate a row without
    the floors and redo it with them when any state of the warp's row comes near one: put cells with (almost) zero and
    negative pressure into the Riemann problem and compare RHS and one RK2 step with the oracle."""
    p, u, weq = setup(oracle, nx, ny, 4)
    rng = np.random.default_rng(3)
    u = u.copy()
    for _ in range(12):
        j, i = int(rng.integers(2, ny - 2)), int(rng.integers(2, nx - 2))
        ke = 0.5 * (u[j, i, 1] ** 2 + u[j, i, 2] ** 2) / u[j, i, 0]
        u[j, i, 3] = ke + rng.choice([0.0, 1e-11, -1e-3, 2e-10]) / (p.gamma - 1.0)      # p = 0, 1e-11, < 0, 2e-10
    dref = oracle.fv2d_compute_update_exact(p, u, weq)
    assert np.all(np.isfinite(dref))
    with wb.FV2D(nx, ny) as s:
        d = s.compute_update_exact(u, weq)
        assert rhs_err(oracle, p, u, d, dref) <= TOL
        got, it, t, dt = s.evolve(u, weq, 1.0, 1)
    ref, it0, t0, dt0, _ = oracle.fv2d_evolve(p, u, weq, 1.0, 1)
    assert it == it0 == 1 and abs(dt - dt0) <= 1e-14 * dt0
    assert rel_linf_fields(got, ref) <= TOL


def test_output_file_of_the_resident_state(wb, oracle, tmp_path):
    """output_file(x,y,u,filen) (benchmark_2d.f90:115-143) from the resident state: x, y, p - p_eq per cell in
    '(7(1PE12.5,1X))', icell outer / jcell inner; packing kernel + asynchronous D2H + host writer thread."""
    nx, ny = 48, 36
    p, u, weq = setup(oracle, nx, ny, 3)
    ref, it, t, dt, cm = oracle.fv2d_evolve(p, u, weq, 1.0, 3)
    with wb.FV2D(nx, ny) as s:
        s.upload(u, weq)
        s.step_async(3)
        s.output_file(str(tmp_path / "fi.dat"), wait=False)      # the writer runs while the next steps are enqueued
        s.step_async(2)
        s.output_wait()
        s.sync()
    lines = (tmp_path / "fi.dat").read_text().splitlines()
    assert len(lines) == nx * ny and all(len(l) == 3 * 12 + 2 for l in lines)
    tab = np.array([[float(v) for v in l.split()] for l in lines]).reshape(nx, ny, 3)
    x, y = oracle.fv2d_get_coords(p)
    w = oracle.fv2d_compute_primitive(p, ref)
    dp = (w[..., 3] - weq[..., 3]).T                      # (nx, ny): the file runs icell outer, jcell inner
    assert np.allclose(tab[..., 0], x.T if x.ndim == 2 else x[:, None], rtol=1e-5, atol=0)
    assert np.abs(tab[..., 2] - dp).max() <= 1e-5 * np.abs(dp).max() + 1e-17
    assert lines[0][:12] == "%12.5E" % float(x.flat[0])


def test_single_rank_has_no_ghost_exchange(wb):
    """wb_fv2d_exchange_kind: 'none' on one rank ('p2p' / 'nccl' on slabs: tools/slab_parity.py exercises both)"""
    with wb.FV2D(64, 48) as s:
        assert s.exchange_kind() == "none"
