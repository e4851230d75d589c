"""CPU tests of the 2D DG oracle (oracle/dg2d.c): known answers implied by 2d/legendre.f90, 2d/test2d.f90,
2d/benchmark_2d_dg.f90 and 2d/limiters.f90."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dg2d.npz")


def smooth_state(o, p, amp=0.2):
    x, y = o.dg2d_get_coords(p)
    rho = 1 + amp * np.sin(2 * np.pi * (x + y)); vx = 1.0 + 0 * x; vy = 0.5 + 0 * x; pr = 1.0 + 0.1 * np.cos(2 * np.pi * x)
    u = np.stack([rho, rho * vx, rho * vy, pr / (p.gamma - 1) + 0.5 * rho * (vx ** 2 + vy ** 2)], axis=-1)
    return x, y, np.ascontiguousarray(u)


def test_gauss_legendre_tables_match_textbook_values(oracle):
    """gl_quadrature (Newton, 500 iterations) must land on the Gauss-Legendre nodes/weights to a few ulp."""
    for m in (1, 2, 3, 4):
        xq, wx, xg, wg = oracle.dg2d_basis(oracle.dg2d_params(mx=m, my=m))
        xr, wr = np.polynomial.legendre.leggauss(m)
        assert np.abs(xq - xr).max() < 4e-16 and np.abs(wx - wr).max() < 1e-15
    # GLL table as shipped (2d/legendre.f90:119-168), incl. the odd weights for n = 2, 3
    _, _, xg, wg = oracle.dg2d_basis(oracle.dg2d_params(mx=2, my=2))
    assert xg.tolist() == [-1.0, 0.0] and wg.tolist() == [1.0, 1.0]
    _, _, xg, wg = oracle.dg2d_basis(oracle.dg2d_params(mx=3, my=3))
    assert xg.tolist() == [-1.0, 0.0, 1.0] and wg.tolist() == [0.75, 0.25, 0.75]


def test_legendre_is_sqrt_2n_plus_1_normalised(oracle):
    for n in range(5):
        assert oracle.dg2d_legendre(1.0, n) == np.sqrt(2.0 * n + 1.0)
        assert oracle.dg2d_legendre(-1.0, n) == (-1) ** n * np.sqrt(2.0 * n + 1.0)
        assert oracle.dg2d_legendre(3.0, n) == oracle.dg2d_legendre(1.0, n)      # argument clamped to [-1,1]
    x = 0.3
    for n in range(1, 5):
        h = 1e-6
        fd = (oracle.dg2d_legendre(x + h, n) - oracle.dg2d_legendre(x - h, n)) / (2 * h)
        assert abs(fd - oracle.dg2d_legendre_prime(x, n)) < 1e-8


@pytest.mark.parametrize("m,bound", [(2, 3e-12), (3, 5e-12)])
def test_test2d_round_trip_fixture(oracle, m, bound):
    """2d/test2d.f90: exp(-x+y) at the GL nodes, project, reconstruct, 1000 round trips; max/min of u-nodes are
    tiny and one-signed: each trip scales the nodes by 0.25*sum(w)^2, which differs from 1 by an ulp or two
    (the sign follows the last bit of the Newton-computed weights, i.e. of libm's cos)."""
    p = oracle.dg2d_params(nx=8, ny=8, mx=m, my=m)
    x, y = oracle.dg2d_get_coords(p)
    u = np.zeros(x.shape + (4,)); u[...] = np.exp(-x + y)[..., None]
    n1 = oracle.dg2d_get_nodes_from_modes(p, oracle.dg2d_get_modes_from_nodes(p, u))
    assert np.abs(n1 - u).max() < 5e-15
    n = n1
    for _ in range(1000):
        n = oracle.dg2d_get_nodes_from_modes(p, oracle.dg2d_get_modes_from_nodes(p, n))
    d = u - n
    _, w, _, _ = oracle.dg2d_basis(p)
    assert 0 < np.abs(d).max() < bound
    if m == 2 and 0.25 * w.sum() ** 2 < 1.0:      # 2 nodes: a pure shrink by 0.25*sum(w)^2 per trip
        assert d.min() > 0


def test_mean_mode_is_the_cell_average(oracle):
    p = oracle.dg2d_params(nx=4, ny=4, mx=3, my=3)
    x, y = oracle.dg2d_get_coords(p)
    u = np.zeros(x.shape + (4,)); u[..., 0] = 2.0 + x * y; u[..., 1] = x ** 2; u[..., 2] = y; u[..., 3] = 1.0
    m = oracle.dg2d_get_modes_from_nodes(p, u)
    dx = 0.25
    xc = (np.arange(4) + 0.5) * dx
    assert np.allclose(m[0, 0, :, :, 2], xc[:, None] * np.ones((1, 4)), atol=1e-15)          # mean of y over cell j
    assert np.allclose(m[0, 0, :, :, 1], (xc ** 2 + dx * dx / 12)[None, :] * np.ones((4, 1)), atol=1e-15)


def test_constant_state_has_zero_rhs(oracle):
    for m in (1, 2, 3):
        p = oracle.dg2d_params(nx=6, ny=6, mx=m, my=m)
        x, y = oracle.dg2d_get_coords(p)
        u = np.zeros(x.shape + (4,)); u[..., 0] = 1.3; u[..., 1] = 0.4; u[..., 2] = -0.2; u[..., 3] = 2.0
        d = oracle.dg2d_compute_update(p, oracle.dg2d_get_modes_from_nodes(p, u), x, y)
        assert np.abs(d).max() < 2e-13


def test_shipped_flux_type_leaves_numerical_flux_zero(oracle):
    """flux_type='llf' (2d/parameters_dg_2d.f90:15) matches no branch of compute_num_flux (:999-1005)."""
    p0 = oracle.dg2d_params(nx=6, ny=6, mx=2, my=2, flux="llf")
    p1 = oracle.dg2d_params(nx=6, ny=6, mx=2, my=2, flux="llf1")
    x, y, u = smooth_state(oracle, p0)
    m = oracle.dg2d_get_modes_from_nodes(p0, u)
    d0, d1 = oracle.dg2d_compute_update(p0, m, x, y), oracle.dg2d_compute_update(p1, m, x, y)
    assert np.abs(d0 - d1).max() > 1e-3
    # with zero numerical flux the mean mode cannot change (volume term of P_0 vanishes, no source)
    assert np.abs(d0[0, 0]).max() == 0.0


def test_periodic_update_conserves_the_mean(oracle):
    p = oracle.dg2d_params(nx=8, ny=8, mx=3, my=3, flux="llf1")
    x, y, u = smooth_state(oracle, p)
    d = oracle.dg2d_compute_update(p, oracle.dg2d_get_modes_from_nodes(p, u), x, y)
    assert np.abs(d[0, 0].sum(axis=(0, 1))).max() < 1e-11


def test_translation_invariance_on_the_periodic_box(oracle):
    p = oracle.dg2d_params(nx=8, ny=8, mx=2, my=2, flux="llf1")
    x, y, u = smooth_state(oracle, p)
    m = oracle.dg2d_get_modes_from_nodes(p, u)
    d = oracle.dg2d_compute_update(p, m, x, y)
    ms = np.ascontiguousarray(np.roll(m, (3, 2), axis=(2, 3)))
    ds = oracle.dg2d_compute_update(p, ms, x, y)
    assert np.array_equal(np.roll(d, (3, 2), axis=(2, 3)), ds)


def test_max_speed_is_order_dependent_like_the_reference(oracle):
    """compute_max_speed (:826-870): velocities of the LAST cell attaining the max speed, cs = min over the cells
    from there to the end of the scan (i outer, j inner)."""
    p = oracle.dg2d_params(nx=4, ny=4, mx=1, my=1)
    m = np.zeros((1, 1, 4, 4, 4)); m[..., 0] = 1.0; m[..., 3] = 2.5
    g = p.gamma
    m[0, 0, 1, 2, :] = [1.0, 3.0, 0.0, 2.5 + 4.5]       # element (i=3, j=2) 1-based: fastest, vx = 3
    m[0, 0, 3, 2, 0] = 4.0                               # later in the scan (i=3, j=4): dense -> small cs
    m[0, 0, 0, 0, 0] = 9.0                               # earlier in the scan: even smaller cs, must NOT count
    cs, vx, vy, sp = oracle.dg2d_compute_max_speed(p, np.ascontiguousarray(m))
    c_fast = np.sqrt(g * (g - 1) * 2.5 / 1.0)
    assert vx == 3.0 and vy == 0.0 and sp == 3.0 + c_fast
    assert cs == np.sqrt(g * ((g - 1) * 2.5) / 4.0)


def test_smooth_convergence_order(oracle):
    """Advected smooth density wave: the mx = 2 scheme converges with order ~2 (the shipped pulse IC is not
    periodic-smooth, so the reference's own 'advection convergence test' stalls at first order in Linf)."""
    errs = []
    for nx in (8, 16):
        p = oracle.dg2d_params(nx=nx, ny=nx, mx=2, my=2, limiter="none", flux="llf1")
        x, y = oracle.dg2d_get_coords(p)
        rho = 1 + 0.2 * np.sin(2 * np.pi * (x + y))
        u = np.ascontiguousarray(np.stack([rho, rho, rho, 1 / (p.gamma - 1) + rho], axis=-1))
        un, it, t, dt = oracle.dg2d_evolve(p, u, x, y, 0.1)
        errs.append(np.abs(un[..., 0] - (1 + 0.2 * np.sin(2 * np.pi * (x + y - 0.2)))).max())
    assert np.log2(errs[0] / errs[1]) > 1.8


def test_ssprk54_weights_drift_as_real4_literals(oracle):
    """the real(4)-rounded convex weights do not sum to 1, so a constant state drifts ~1e-8 per step."""
    p = oracle.dg2d_params(nx=4, ny=4, mx=2, my=2, limiter="none", flux="llf1")
    x, y = oracle.dg2d_get_coords(p)
    u = np.zeros(x.shape + (4,)); u[..., 0] = 1.0; u[..., 3] = 2.5
    un, it, t, dt = oracle.dg2d_evolve(p, u, x, y, 1.0, 10)
    drift = un[..., 0].mean() - 1.0
    assert it == 10 and -5e-7 < drift < -1e-8


def test_positivity_limiter_keeps_means_and_restores_positivity(oracle):
    p = oracle.dg2d_params(nx=6, ny=6, mx=3, my=3, limiter="ONP")
    rng = np.random.default_rng(3)
    m = np.zeros((3, 3, 6, 6, 4)); m[0, 0, ..., 0] = 1.0; m[0, 0, ..., 3] = 2.5
    m[1:, :, ..., 0] = 0.6 * rng.standard_normal((2, 3, 6, 6)); m[0, 1:, ..., 0] = 0.6 * rng.standard_normal((2, 6, 6))
    m[..., 1] = 0.3 * rng.standard_normal((3, 3, 6, 6)); m[1, 1, ..., 3] = 1.5 * rng.standard_normal((6, 6))
    lim = oracle.dg2d_apply_limiter(p, np.ascontiguousarray(m))
    assert np.array_equal(lim[0, 0], m[0, 0])                      # cell means untouched
    nodes = oracle.dg2d_get_nodes_from_modes(p, lim)
    assert nodes[..., 0].min() > 0
    assert np.abs(lim).sum() < np.abs(m).sum()                     # something was limited
    # idempotent up to round-off: a second pass changes (almost) nothing
    lim2 = oracle.dg2d_apply_limiter(p, lim)
    assert np.abs(lim2 - lim).max() < 1e-12


def test_hio_and_1or_leave_linear_data_alone_and_clip_oscillations(oracle):
    for lim in ("HIO", "1OR", "LOW"):
        p = oracle.dg2d_params(nx=8, ny=8, mx=3, my=3, limiter=lim, bc=2)
        x, y = oracle.dg2d_get_coords(p)
        u = np.zeros(x.shape + (4,)); u[..., 0] = 1.0 + 0.5 * x + 0.25 * y; u[..., 3] = 2.5
        m = oracle.dg2d_get_modes_from_nodes(p, u)
        out = oracle.dg2d_apply_limiter(p, m)
        if lim == "HIO":
            assert np.abs(out[..., 1:-1, 1:-1, :] - m[..., 1:-1, 1:-1, :]).max() < 1e-14   # interior untouched
        if lim == "LOW":
            assert np.all(out[0, 1:] == 0) and np.all(out[1:, 0] == 0) and np.array_equal(out[0, 0], m[0, 0])
        if lim == "1OR":     # a saw-tooth in the linear x-mode of the density is clipped against the neighbours' means
            m2 = m.copy(); m2[0, 1, :, ::2, 0] += 0.3
            out2 = oracle.dg2d_apply_limiter(p, np.ascontiguousarray(m2))
            assert np.abs(out2[0, 1, :, 2:-2:2, 0]).max() < 0.5 * np.abs(m2[0, 1, :, 2:-2:2, 0]).max()
        if lim == "HIO":     # the hierarchy starts at the highest diagonal mode: an oscillating (3,3) mode is removed
            m2 = m.copy(); m2[2, 2, :, ::2, 0] += 0.3
            out2 = oracle.dg2d_apply_limiter(p, np.ascontiguousarray(m2))
            assert np.abs(out2[2, 2, 2:-2, 2:-2, 0]).max() < 0.1 * np.abs(m2[2, 2, 2:-2, 2:-2, 0]).max()


def test_golden_vectors(oracle):
    g = np.load(GOLD)
    for tag in [k[:-5] for k in g.files if k.endswith("_meta")]:
        nx, mx, bc, source, gcase, flux, lim, solver, ninit, steps = (int(v) for v in g[f"{tag}_meta"])
        p = oracle.dg2d_params(nx=nx, ny=nx, mx=mx, my=mx, bc=bc, source=source, grad_phi_case=gcase, ninit=ninit)
        p.flux_id, p.limiter_id, p.solver_id = flux, lim, solver
        x, y = oracle.dg2d_get_coords(p)
        u0 = g[f"{tag}_u0"]
        m0 = oracle.dg2d_get_modes_from_nodes(p, u0)
        assert np.array_equal(oracle.dg2d_compute_update(p, m0, x, y), g[f"{tag}_dudt"])
        assert np.array_equal(oracle.dg2d_apply_limiter(p, m0), g[f"{tag}_lim"])
        un, it, t, dt = oracle.dg2d_evolve(p, u0, x, y, 1.0, steps)
        assert np.array_equal(un, g[f"{tag}_un"]) and np.array_equal(np.array([it, t, dt]), g[f"{tag}_clock"])


# ---------------------------------------------------------------- 'hll2' / 'hllc' (compute_hllflux :1008-1026, compute_hllcflux :1030-1134)
def _phys_flux(u, flag, gamma=float(np.float32(1.4))):
    rho, mx, my, E = u
    vx, vy = mx / rho, my / rho
    p = (gamma - 1.0) * (E - 0.5 * rho * (vx * vx + vy * vy))
    return np.array([mx, mx * vx + p, mx * vy, vx * (E + p)]) if flag == 1 else np.array([my, mx * vy, my * vy + p, vy * (E + p)])


def test_hll2_as_shipped_is_consistent_and_left_upwind_for_subsonic_states(oracle):
    """compute_hllflux (:1008-1026) builds its wave speeds from the isotropic |v| and c: a_plus = max(0, c + |v|) and
    a_minus = max(0, -(c - |v|)) = max(0, |v| - c).  For subsonic states a_minus is therefore ZERO and the formula
    collapses to f_left, whatever the right state; only when some |v| exceeds c does the right state enter."""
    p = oracle.dg2d_params(flux="hll2")
    u = np.array([1.3, 0.4, -0.2, 2.0])
    for flag in (1, 2):
        assert np.allclose(oracle.dg2d_num_flux(p, u, u, flag), _phys_flux(u, flag), rtol=1e-14)
    ul = np.array([1.3, 0.4, -0.2, 2.0]); ur = np.array([0.9, -0.3, 0.1, 1.7])                    # both subsonic
    for flag in (1, 2):
        assert np.allclose(oracle.dg2d_num_flux(p, ul, ur, flag), _phys_flux(ul, flag), rtol=1e-14)
    ul = np.array([1.0, 3.0, 0.0, 7.0]); ur = np.array([0.5, 1.6, 0.0, 3.31])                      # |v| > c on both sides
    g = float(np.float32(1.4))
    sp = lambda q: (abs(q[1] / q[0]), np.sqrt(g * (g - 1) * (q[3] - 0.5 * q[1] ** 2 / q[0]) / q[0]))
    (vl, cl), (vr, cr) = sp(ul), sp(ur)
    ap, am = max(cl + vl, cr + vr), max(vl - cl, vr - cr)
    expect = (ap * _phys_flux(ul, 1) + am * _phys_flux(ur, 1) - ap * am * (ur - ul)) / (ap + am)
    assert am > 0 and np.allclose(oracle.dg2d_num_flux(p, ul, ur, 1), expect, rtol=1e-14)


def test_hllc_as_shipped_consistent_only_where_its_typos_are_silent(oracle):
    """Equal left and right states: S_M = v_n and the left star state is the state itself, so the branches that use it
    (v_n > 0) return the physical flux.  The right star ENERGY carries a misplaced parenthesis (:1075, :1116):
    E* - E = rho (v_n - v_n (v_n + p / (rho c))) != 0, so for -c <= v_n <= 0 the energy flux is off by S_R times that --
    except at v_n = 0, where it vanishes.  The y branch also takes the x momentum of the right star state from the LEFT
    velocity (:1114): silent for equal states."""
    p = oracle.dg2d_params(flux="hllc")
    g = float(np.float32(1.4))
    for u, flag in ((np.array([1.3, 0.4, -0.2, 2.0]), 1), (np.array([1.3, 0.4, 0.2, 2.0]), 2), (np.array([1.3, 0.4, 0.0, 2.0]), 2)):
        assert np.allclose(oracle.dg2d_num_flux(p, u, u, flag), _phys_flux(u, flag), rtol=1e-13, atol=1e-15)
    u = np.array([1.3, 0.4, -0.2, 2.0])                               # y face, v_y = -0.154 (subsonic, negative)
    rho, vx, vy = u[0], u[1] / u[0], u[2] / u[0]
    pr = (g - 1.0) * (u[3] - 0.5 * rho * (vx * vx + vy * vy)); c = np.sqrt(g * pr / rho)
    SR = vy + c
    dE = rho * (vy - vy * (vy + pr / (rho * (SR - vy)))) 
    expect = _phys_flux(u, 2) + SR * np.array([0.0, 0.0, 0.0, dE])
    assert np.allclose(oracle.dg2d_num_flux(p, u, u, 2), expect, rtol=1e-13)
    # :1114 -- different x velocities on the two sides of a y face with S_M <= 0 <= S_R: the x momentum of the right star
    # state is rho* times the LEFT x velocity
    ul = np.array([1.0, 0.5, -0.1, 2.5]); ur = np.array([1.0, -0.3, -0.1, 2.5])
    f = oracle.dg2d_num_flux(p, ul, ur, 2)
    # momentum-x flux = f2(ur)[1] + S_R (rho*_R vx_L - mx_R)
    wl = ul[1] / ul[0]; vyr = ur[2] / ur[0]
    prl = (g - 1) * (ul[3] - 0.5 * (ul[1] ** 2 + ul[2] ** 2) / ul[0]); prr = (g - 1) * (ur[3] - 0.5 * (ur[1] ** 2 + ur[2] ** 2) / ur[0])
    cl, cr = np.sqrt(g * prl / ul[0]), np.sqrt(g * prr / ur[0])
    SL, SR = min(-0.1, vyr) - max(cl, cr), max(-0.1, vyr) + max(cl, cr)
    SM = (ur[0] * vyr * (SR - vyr) - ul[0] * (-0.1) * (SL + 0.1) + prl - prr) / (ur[0] * (SR - vyr) - ul[0] * (SL + 0.1))
    rs = ur[0] * (SR - vyr) / (SR - SM)
    assert SM <= 0 <= SR and np.isclose(f[1], ur[1] * vyr + SR * (rs * wl - ur[1]), rtol=1e-13)


@pytest.mark.parametrize("flux", ["hll2", "hllc"])
def test_hll_fluxes_keep_the_scheme_conservative_and_convergent(oracle, flux):
    """Whatever its formula, a numerical flux that is single valued per face point conserves the mean on the periodic box;
    and for a right-moving smooth state (every branch the typos spare) both give the llf1 RHS up to the dissipation."""
    p = oracle.dg2d_params(nx=8, ny=8, mx=3, my=3, flux=flux)
    x, y, u = smooth_state(oracle, p)
    m = oracle.dg2d_get_modes_from_nodes(p, u)
    d = oracle.dg2d_compute_update(p, m, x, y)
    assert np.all(np.isfinite(d)) and np.abs(d[0, 0].sum(axis=(0, 1))).max() < 1e-11
    d1 = oracle.dg2d_compute_update(oracle.dg2d_params(nx=8, ny=8, mx=3, my=3, flux="llf1"), m, x, y)
    assert 0 < np.abs(d - d1).max() < 0.2 * np.abs(d1).max()
