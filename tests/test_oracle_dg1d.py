"""CPU tests of the 1D DG oracle (oracle/dg1d.c): dg_with_source.f90 default path ('RKi') + root legendre.f90."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dg1d.npz")
f32 = lambda v: float(np.float32(v))  # noqa: E731


def test_root_legendre_uses_single_precision_constants(oracle):
    """P0 = 0.7071067690849304, P1 factor 1.2247449159622192, P2 factor 0.7905694246292114;
    GL-3 nodes -/+0.774596631526947, weights 0.5555555820465088 / 0.8888888955116272; GL-2 nodes -/+0.5773502588272095."""
    assert oracle.dg1d_legendre(0.3, 0) == 0.7071067690849304
    assert oracle.dg1d_legendre(1.0, 1) == 1.2247449159622192
    assert oracle.dg1d_legendre(1.0, 2) == 0.25 * 2.0 * float(np.sqrt(np.float32(10.0)))
    x, w = oracle.dg1d_quadrature(oracle.dg1d_params(n=3))
    assert x.tolist() == [-0.774596631526947, 0.0, 0.774596631526947]
    assert w.tolist() == [0.5555555820465088, 0.8888888955116272, 0.5555555820465088]
    x, w = oracle.dg1d_quadrature(oracle.dg1d_params(n=2))
    assert x.tolist() == [-0.5773502588272095, 0.5773502588272095] and w.tolist() == [1.0, 1.0]


def test_steady_state_rhs_llf_exact_hllc_ulp_over_dx(oracle):
    """ninit=7 -> delta_u == 0; the RHS is exactly zero with riemann_llf and <= 4 ulp(p)/dx with the default
    HLLC (its star-state algebra differs from the physical flux by an ulp on some faces)."""
    for n in (1, 2, 3):
        for nx in (64, 128):
            p = oracle.dg1d_params(n=n, nx=nx, riemann=1, ninit=7)
            ui, ueq, du = oracle.dg1d_setup(p)
            assert np.all(du == 0.0)
            assert np.all(oracle.dg1d_compute_update_exact_delta(p, du, ueq) == 0.0)
            p = oracle.dg1d_params(n=n, nx=nx, riemann=2, ninit=7)
            d = oracle.dg1d_compute_update_exact_delta(p, du, ueq)
            assert np.abs(d).max() <= 4 * 2.220446049250313e-16 * nx * 1.3      # P(+-1) up to 1.58 for n = 3
    p = oracle.dg1d_params(riemann=1, ninit=7)
    ui, ueq, du = oracle.dg1d_setup(p)
    du2, ui2, it, t, dt = oracle.dg1d_evolve_rki(p, du, ueq, ui, 0.2)
    assert it > 100 and np.all(du2 == 0.0) and np.array_equal(ui2, ueq)


def test_end_cells_are_frozen_and_time_step_formula(oracle):
    p = oracle.dg1d_params(ninit=8, pert=1e-3)
    ui, ueq, du = oracle.dg1d_setup(p)
    d = oracle.dg1d_compute_update_exact_delta(p, du, ueq)
    assert np.all(d[0] == 0) and np.all(d[-1] == 0) and np.abs(d[1:-1]).max() > 0
    c = oracle.dg1d_compute_max_speed(p, ui)
    _, _, it, t, dt = oracle.dg1d_evolve_rki(p, du, ueq, ui, 1.0, 1)
    assert dt == f32(0.9) * (1.0 / 128) / c / 7.0 and it == 1 and t == dt


def test_projection_carries_the_half_factor(oracle):
    """orthonormal P0..P2 but Teyssier's 0.5 projection factor (:161) -> the mean mode is
    0.5*sqrt(0.5)*integral, i.e. half of the orthonormal-basis coefficient."""
    p = oracle.dg1d_params(n=3, nx=16, ninit=8, pert=1.0)
    ui, ueq, du = oracle.dg1d_setup(p)
    x, w = oracle.dg1d_quadrature(p)
    diff = ui - ueq
    mean_mode = 0.5 * (diff * w[None, :, None]).sum(axis=1) * 0.7071067690849304
    assert np.allclose(du[:, 0, :], mean_mode, rtol=1e-14, atol=1e-18)


def test_perturbation_evolves_linearly_in_its_amplitude(oracle):
    outs = []
    for pert in (1e-6, 1e-8):
        p = oracle.dg1d_params(nx=64, ninit=8, pert=pert)
        ui, ueq, du = oracle.dg1d_setup(p)
        du2, ui2, it, t, dt = oracle.dg1d_evolve_rki(p, du, ueq, ui, 0.05)
        outs.append(du2 / pert)
    assert np.abs(outs[0] - outs[1]).max() < 1e-4 * np.abs(outs[0]).max()


def test_golden_vectors(oracle):
    g = np.load(GOLD)
    for tag in [k[:-5] for k in g.files if k.endswith("_meta")]:
        n, nx, riemann, source, steps = (int(v) for v in g[f"{tag}_meta"])
        p = oracle.dg1d_params(n=n, nx=nx, riemann=riemann, source=source)
        du, ueq, ui = g[f"{tag}_du"], g[f"{tag}_ueq"], g[f"{tag}_ui"]
        assert np.array_equal(oracle.dg1d_compute_update_exact_delta(p, du, ueq), g[f"{tag}_dudt"])
        du2, ui2, it, t, dt = oracle.dg1d_evolve_rki(p, du, ueq, ui, 1.0, steps)
        assert np.array_equal(du2, g[f"{tag}_du2"]) and np.array_equal(ui2, g[f"{tag}_ui2"])
        assert np.array_equal(np.array([it, t, dt]), g[f"{tag}_clock"])


# ---------------------------------------------------------------- plain update / limiter / 'RK1'..'RK4' (:807-1028, :414-519, :173-230)
def _smooth_periodic_modes(oracle, p):
    """Projection (program dg :33-47, 0.5 factor) of a smooth periodic full state sampled at the quadrature nodes."""
    x, w = oracle.dg1d_quadrature(p)
    dx = p.boxlen / p.nx
    xc = (np.arange(p.nx) + 0.5) * dx
    xq = xc[:, None] + 0.5 * dx * x[None, :]
    rho = 1.0 + 0.2 * np.sin(2 * np.pi * xq)
    v = 0.3 + 0.0 * xq
    pr = 1.0 + 0.0 * xq
    un = np.stack([rho, rho * v, pr / (p.gamma - 1.0) + 0.5 * rho * v * v], axis=-1)
    return oracle.dg1d_project(p, un), un


def test_plain_update_constant_state_and_end_cell_copies(oracle):
    """compute_update on a constant state without source: every face flux equals the physical flux, so the mean mode
    of dudt is round-off; for the higher modes the volume integral of a constant flux against P'_i only cancels the
    edge terms up to the real(4) GL weights of the root legendre.f90 (sum = 2 + 6e-8) -> O(6e-8 F/dx),
    a property of the reference; :1026-1027 copy cell 2 -> 1 and cell nx -> nx-1."""
    for riemann in (1, 2):
        p = oracle.dg1d_params(n=3, nx=32, riemann=riemann, source=1, bc=1)
        un = np.broadcast_to(np.array([1.3, 0.4, 2.2]), (p.nx, p.n, 3)).copy()
        u = oracle.dg1d_project(p, un)
        d = oracle.dg1d_compute_update(p, u)
        assert np.abs(d[:, 0, :]).max() <= 1e-12
        assert 1e-7 < np.abs(d[:, 1:, :]).max() <= 6e-8 * 5.0 * p.nx * 2
    p = oracle.dg1d_params(n=3, nx=32, riemann=2, source=2, bc=1)
    u, _ = _smooth_periodic_modes(oracle, p)
    d = oracle.dg1d_compute_update(p, u)
    assert np.array_equal(d[0], d[1]) and np.array_equal(d[-2], d[-1]) and not np.array_equal(d[2], d[3])


def test_plain_update_is_consistent_with_the_euler_equations(oracle):
    """Mean mode of dudt ~ -(F(x+)-F(x-))/dx * P0-projection: for pure advection of a density wave (v, p constant)
    d(rho)/dt = -v d(rho)/dx.  SURVEY 9.10: with the orthonormal root basis, the 0.5 projection factor and the 1/dx
    (not 2/dx) in :1019 the scheme is a consistent DG running at HALF speed (the Euler flux is homogeneous of degree 1,
    so the halved reconstructed state gives a halved flux): the cell-mean rate is 0.5 x the analytic one."""
    errs = []
    for nx in (32, 64):
        p = oracle.dg1d_params(n=3, nx=nx, riemann=1, source=1, bc=1)
        u, un = _smooth_periodic_modes(oracle, p)
        d = oracle.dg1d_compute_update(p, u)
        dx = 1.0 / nx
        xc = (np.arange(nx) + 0.5) * dx
        # cell average of -v*rho_x = -0.3*0.2*(sin(2pi x+) - sin(2pi x-))/dx
        exact = -0.3 * 0.2 * (np.sin(2 * np.pi * (xc + dx / 2)) - np.sin(2 * np.pi * (xc - dx / 2))) / dx
        # mean mode = 0.5 * P0 * integral over [-1,1] = P0 * cell average (P0 = sqrt(.5) f32)
        got = d[2:-2, 0, 0] / 0.7071067690849304
        errs.append(np.abs(got - 0.5 * exact[2:-2]).max())
    assert errs[0] < 1e-3 and errs[1] < errs[0] / 7          # ~ third-order decay of the mean-rate error


def test_limiter_keeps_means_and_smooth_data_and_flattens_negative_cells(oracle):
    p = oracle.dg1d_params(n=3, nx=64, bc=1, use_limiter=1)
    u, _ = _smooth_periodic_modes(oracle, p)
    v = oracle.dg1d_limiter(p, u)
    assert np.array_equal(v[:, 0, :], u[:, 0, :])                               # means untouched
    # smooth data: the slopes survive (characteristic round trip only); the curvature moment is clipped where it
    # changes sign (minmod of neighbour slope differences) by at most its own size, and the 1 % test (:478) then
    # stops the cascade before it reaches the slope
    assert np.abs(v[:, 1] - u[:, 1]).max() <= 1e-15
    assert 0 < np.abs(v[:, 2] - u[:, 2]).max() <= 0.06 * np.abs(u[:, 2]).max()
    p0 = oracle.dg1d_params(n=3, nx=64, bc=1, use_limiter=0)
    assert np.array_equal(oracle.dg1d_limiter(p0, u), u)                        # positivity part alone: identity here
    bad = u.copy()
    bad[10, 1, 0] = 5.0                                                         # slope that drives rho(-1) negative
    w = oracle.dg1d_limiter(p0, bad)
    assert np.all(w[10, 1:, :] == 0.0) and np.array_equal(w[10, 0], bad[10, 0])
    assert np.array_equal(np.delete(w, 10, axis=0), np.delete(bad, 10, axis=0))
    p1 = oracle.dg1d_params(n=1, nx=8, bc=1, use_limiter=1)
    u1 = np.random.default_rng(0).random((8, 1, 3)) + 1
    assert np.array_equal(oracle.dg1d_limiter(p1, u1), u1)                      # n == 1: return (:431)


def test_limiter_on_a_jump_is_tvd_in_the_means_sense(oracle):
    """Square density pulse projected on order-3 modes: the moment limiter zeroes the higher moments of the cells next
    to the jump (minmod of differences with opposite signs) and leaves the flat regions alone."""
    p = oracle.dg1d_params(n=3, nx=40, bc=2, use_limiter=1)
    un = np.zeros((40, 3, 3))
    un[..., 0] = 1.0; un[..., 2] = 2.5
    un[15:25, :, 0] = 2.0
    u = oracle.dg1d_project(p, un)
    u[15, 1, 0] = 0.3; u[24, 1, 0] = -0.3                                        # spurious slopes at the extrema
    v = oracle.dg1d_limiter(p, u)
    assert abs(v[15, 1, 0]) <= 1e-15 + abs(u[15, 1, 0]) and np.abs(v[5, 1:, :]).max() == 0.0
    assert np.abs(v[24, 1, 0]) <= abs(u[24, 1, 0])


@pytest.mark.parametrize("integ,order", [("RK2", 2), ("RK3", 3)])
def test_rk_paths_advect_a_density_wave(oracle, integ, order):
    """'RK2'/'RK3' on the plain update (periodic, no source): a density wave moves with v = 0.3 at half speed
    (SURVEY 9.10); the cell means after t = 0.05 match the shifted profile with an error that decays at least ~4x per refinement (time step ~ dx)."""
    errs = []
    for nx in (16, 32):
        p = oracle.dg1d_params(n=3, nx=nx, riemann=1, source=1, bc=1)
        u, un = _smooth_periodic_modes(oracle, p)
        du = np.zeros_like(u)
        u2, ui2, it, t, dt = oracle.dg1d_evolve_rk(p, integ, u, du, un, un, 0.05)
        dx = 1.0 / nx
        xc = (np.arange(nx) + 0.5) * dx
        s = 0.5 * 0.3 * t
        mean = 1.0 - 0.2 * (np.cos(2 * np.pi * (xc + dx / 2 - s)) - np.cos(2 * np.pi * (xc - dx / 2 - s))) / (2 * np.pi * dx)
        got = u2[3:-3, 0, 0] / 0.7071067690849304
        errs.append(np.abs(got - mean[3:-3]).max())
        assert t >= 0.05 and it > 3
    assert errs[1] < errs[0] / 4 and errs[1] < 2e-4


def test_rk4_path_subtracts_the_nodal_equilibrium_as_shipped(oracle):
    """'RK4' (:205-227) does u = u - u_eq (modes minus NODAL values, as shipped) before the stages and adds it back:
    with u_eq == 0 it must coincide with running the five stages directly; with a non-zero u_eq it differs."""
    p = oracle.dg1d_params(n=3, nx=24, riemann=2, source=1, bc=1)
    u, un = _smooth_periodic_modes(oracle, p)
    z = np.zeros_like(u)
    a, _, it, t, dt = oracle.dg1d_evolve_rk(p, "RK4", u, z, z, un, 1.0, 1)
    b, _, _, _, _ = oracle.dg1d_evolve_rk(p, "RK3", u, z, z, un, 1.0, 1)
    assert it == 1 and np.abs(a - b).max() < 1e-6 and np.abs(a - b).max() > 0
    c, _, _, _, _ = oracle.dg1d_evolve_rk(p, "RK4", u, z, 0.01 * un, un, 1.0, 1)
    assert np.abs(c - a).max() > 1e-8


def test_golden_vectors_plain_paths(oracle):
    g = np.load(GOLD)
    tags = [k[:-6] for k in g.files if k.endswith("_pmeta")]
    assert tags
    for tag in tags:
        n, nx, riemann, source, bc, use_limiter, integ, steps = (int(v) for v in g[f"{tag}_pmeta"])
        p = oracle.dg1d_params(n=n, nx=nx, riemann=riemann, source=source, bc=bc, use_limiter=use_limiter)
        u, du, ueq, ui = g[f"{tag}_u"], g[f"{tag}_du"], g[f"{tag}_ueq"], g[f"{tag}_ui"]
        assert np.array_equal(oracle.dg1d_compute_update(p, u), g[f"{tag}_dudt"]), tag
        assert np.array_equal(oracle.dg1d_limiter(p, u), g[f"{tag}_lim"]), tag
        u2, ui2, it, t, dt = oracle.dg1d_evolve_rk(p, f"RK{integ}", u, du, ueq, ui, 1.0, steps)
        assert np.array_equal(u2, g[f"{tag}_u2"]) and np.array_equal(ui2, g[f"{tag}_ui2"]), tag
        assert np.array_equal(np.array([it, t, dt]), g[f"{tag}_pclock"]), tag


# ---------------------------------------------------------------- 'RKw' / 'RKe': compute_update_exact, limiter_TDV, limiter_cons
def _w_setup(oracle, **kw):
    p = oracle.dg1d_params(**kw)
    ui, ueq, du = oracle.dg1d_setup(p)
    return p, ui, ueq, du, oracle.dg1d_project(p, ui), oracle.dg1d_project(p, ueq)


def test_update_exact_interior_is_independent_of_the_end_treatment(oracle):
    """compute_update_exact (:1380-1744): bc 4 and bc 5 differ only in the boundary states that replace the out-of-bounds
    Riemann problems of faces 1 and nx+1 (:1516-1620) -> every cell but the two end cells gets the same dudt, bit for bit."""
    out = {}
    for bc in (4, 5):
        p, ui, ueq, du, u, q = _w_setup(oracle, n=3, nx=48, riemann=2, source=2, ninit=8, pert=1e-3, bc=bc)
        out[bc] = oracle.dg1d_compute_update_exact(p, u, q)
        assert np.all(np.isfinite(out[bc]))
    assert np.array_equal(out[4][1:-1], out[5][1:-1])
    assert not np.array_equal(out[4][0], out[5][0]) and not np.array_equal(out[4][-1], out[5][-1])


def test_update_exact_volume_and_source_terms_cancel_at_u_equal_u_eq(oracle):
    """With u == u_eq_modes the volume and source integrals of the state and of the equilibrium are the same numbers, so
    dudt is the face part alone.  As shipped that part does NOT vanish: the modes carry Teyssier's 0.5 projection factor
    (traces = half the state) while flux_face_eq is the flux of the full equilibrium (:1411-1416), so the
    mean mode of an interior cell is left with (1/dx) * P0 * 0.5 * (p(x_right) - p(x_left)) ~ -0.5 P0 exp(-x):
    'RKw' is not well balanced in the reference, which is why 'RKi' is its default."""
    p, ui, ueq, du, u, q = _w_setup(oracle, n=3, nx=64, riemann=1, source=2, ninit=7, bc=5)
    d = oracle.dg1d_compute_update_exact(p, q, q)
    dx = 1.0 / 64
    xl = np.arange(64) * dx
    expect = 0.5 * 0.7071067690849304 * (np.exp(-(xl + dx)) - np.exp(-xl)) / dx
    assert np.allclose(d[1:-1, 0, 1], expect[1:-1], rtol=2e-6)            # momentum equation, mean mode
    assert np.abs(d[1:-1, 0, 0]).max() < 1e-6 * np.abs(d[1:-1, 0, 1]).max() + 1e-7   # mass: only LLF dissipation of the jump


def test_limiter_tdv_and_cons_positivity_parts(oracle):
    p = oracle.dg1d_params(n=3, nx=32, bc=5, use_limiter=0)
    rng = np.random.default_rng(4)
    u = np.zeros((32, 3, 3)); u[:, 0, 0] = 1.0; u[:, 0, 2] = 2.5
    u[:, 1:, :] = 0.01 * rng.standard_normal((32, 2, 3))
    assert np.array_equal(oracle.dg1d_limiter_tdv(p, u), u)                      # positive traces: untouched
    bad = u.copy(); bad[7, 1, 0] = 5.0; bad[20, 1, 2] = 30.0                      # negative density / energy trace on the left
    v = oracle.dg1d_limiter_tdv(p, bad)
    for c in (7, 20):
        assert np.all(v[c, 1:] == 0) and np.array_equal(v[c, 0], bad[c, 0])
    assert np.array_equal(np.delete(v, (7, 20), axis=0), np.delete(bad, (7, 20), axis=0))
    # limiter_cons without the moment part is the positivity fallback of limiter() (the same lines, :480-517 / :715-732)
    assert np.array_equal(oracle.dg1d_limiter_cons(p, bad), oracle.dg1d_limiter(p, bad))
    p1 = oracle.dg1d_params(n=3, nx=32, bc=2, use_limiter=1)
    w = oracle.dg1d_limiter_cons(p1, u)
    assert np.array_equal(w[:, 0], u[:, 0]) and np.all(np.abs(w[:, 1:]) <= np.abs(u[:, 1:]) + 1e-300)   # minmod never grows a moment


def test_rke_keeps_the_discrete_steady_state_and_rkw_runs_as_shipped(oracle):
    p, ui, ueq, du, u, q = _w_setup(oracle, n=3, nx=64, riemann=1, source=2, ninit=7, bc=5)
    uu, dd, ui2, it, t, dt = oracle.dg1d_evolve_w(p, "RKe", u, du, ueq, q, ui, 0.05)
    assert it > 20 and np.all(dd == 0.0) and np.array_equal(ui2, ueq)            # delta-form + LLF: exactly steady
    p, ui, ueq, du, u, q = _w_setup(oracle, n=3, nx=64, riemann=2, source=2, ninit=8, pert=1e-3, bc=5)
    c = oracle.dg1d_compute_max_speed(p, ui)
    uu, dd, ui2, it, t, dt = oracle.dg1d_evolve_w(p, "RKw", u, du, ueq, q, ui, 1.0, 1)
    assert it == 1 and dt == f32(0.9) * (1.0 / 64) / c / 7.0 and np.all(np.isfinite(uu))
    # the limiter acts on u - u_eq_modes: what comes back as delta_u is exactly that difference or its flattened version
    diff = uu - q
    same = np.isclose(dd, diff, rtol=0, atol=1e-15).all(axis=(1, 2))
    flat = (dd[:, 1:] == 0).all(axis=(1, 2))
    assert np.all(same | flat)


def test_golden_vectors_w_paths(oracle):
    g = np.load(GOLD)
    tags = [k[:-6] for k in g.files if k.endswith("_wmeta")]
    assert tags
    for tag in tags:
        n, nx, riemann, source, bc, use_limiter, integ, steps = (int(v) for v in g[f"{tag}_wmeta"])
        p = oracle.dg1d_params(n=n, nx=nx, riemann=riemann, source=source, bc=bc, use_limiter=use_limiter)
        u, du, ueq, q, ui = g[f"{tag}_u"], g[f"{tag}_du"], g[f"{tag}_ueq"], g[f"{tag}_q"], g[f"{tag}_ui"]
        if bc in (4, 5):
            assert np.array_equal(oracle.dg1d_compute_update_exact(p, u, q), g[f"{tag}_dudt"]), tag
        assert np.array_equal(oracle.dg1d_limiter_cons(p, g[f"{tag}_lin"]), g[f"{tag}_lcons"]), tag
        if not use_limiter:
            assert np.array_equal(oracle.dg1d_limiter_tdv(p, g[f"{tag}_lin"]), g[f"{tag}_ltdv"]), tag
        uu, dd, ui2, it, t, dt = oracle.dg1d_evolve_w(p, "RKw" if integ == 5 else "RKe", u, du, ueq, q, ui, 1.0, steps)
        assert np.array_equal(uu, g[f"{tag}_u2"]) and np.array_equal(dd, g[f"{tag}_du2"]) and np.array_equal(ui2, g[f"{tag}_ui2"]), tag
        assert np.array_equal(np.array([it, t, dt]), g[f"{tag}_wclock"]), tag
