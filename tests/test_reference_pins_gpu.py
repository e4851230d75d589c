"""GPU parity of the CUDA path (through the C-ABI) DIRECTLY against vectors produced by the reference's own source text
(tests/golden/ref_*.npz, made by tests/golden/make_ref_golden.py: the reference .f90 files executed by
oracle/f90interp.py).  No oracle in the loop: inputs and expected outputs both come from the committed files.

Bars: BIT-FOR-BIT for the reference-order kernels wherever no transcendental is evaluated on the device (CUDA's
exp/pow differ from glibc's in the last bit); 1e-12 relative L-inf (north star) for the fused production kernels and
for paths that evaluate exp/pow on the device."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-12


def gold(name):
    return np.load(os.path.join(HERE, "golden", name))


def tags(name, prefix=""):
    return sorted({k.split("/")[0] for k in gold(name).files if k.startswith(prefix)})


@pytest.fixture(scope="module")
def wb():
    import __graft_entry__ as ge
    ge.build()
    import wbeuler
    return wbeuler


def rel(a, b):
    b = np.asarray(b)
    m = ~np.isnan(b)
    den = np.abs(b[m]).max()
    num = np.abs(np.asarray(a)[m] - b[m]).max()
    return num / den if den > 0 else num


def field_err(a, b):
    """per conserved field, relative to the norm of the whole state (fields that are ~0 carry only rounding noise)"""
    return np.abs(a - b).max() / np.abs(b).max()


# ------------------------------------------------------------------------------------------------ 2D FV
@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("tag", tags("ref_fv2d.npz"))
def test_fv2d_cuda_equals_reference_source(wb, tag, arith):
    g = gold("ref_fv2d.npz")
    nx, ny, ninit, neq, steps = (int(v) for v in g[f"{tag}/meta"])
    u, weq = g[f"{tag}/u"], g[f"{tag}/weq"]
    with wb.FV2D(nx, ny, nequilibrium=neq, arith=arith, device=0) as s:
        cmax = s.compute_max_speed(u)
        assert abs(cmax - float(g[f"{tag}/cmax"])) <= 4e-16 * cmax
        d = s.compute_update_exact(u, weq)
        dref = g[f"{tag}/dudt"]
        if not np.any(dref):
            assert not np.any(d), "hydrostatic state: the RHS must be bitwise zero like the reference's"
        dt = 0.5 * (1.0 / nx) / cmax * 0.5
        assert np.abs(dt * (d - dref)).max() / np.abs(u).max() <= TOL
        if arith == 1 and f"{tag}/dudt_plain" in g.files:
            dp = s.compute_update(u, weq)
            assert np.abs(dt * (dp - g[f"{tag}/dudt_plain"])).max() / np.abs(u).max() <= TOL
        un, it, t, dtl = s.evolve(u, weq, float(g[f"{tag}/tend"]), -1)
        assert it == steps
        assert field_err(un, g[f"{tag}/u_evolved"]) <= TOL


# ------------------------------------------------------------------------------------------------ 2D DG
def _dg2d(wb, g, tag, arith):
    n, m, bc, source, gcase, ninit, steps = (int(v) for v in g[f"{tag}/meta"])
    flux, lim, solver = (str(s) for s in g[f"{tag}/names"])
    box = float(g[f"{tag}/boxlen"]) if f"{tag}/boxlen" in g.files else 1.0
    return wb.DG2D(nx=n, ny=n, mx=m, my=m, bc=bc, source=source, grad_phi_case=gcase, flux=flux, limiter=lim, solver=solver,
                   ninit=ninit, device=0, arith=arith, boxlen_x=box, boxlen_y=box), steps, gcase, source


@pytest.mark.parametrize("tag", tags("ref_dg2d.npz"))
def test_dg2d_reference_order_kernels_equal_reference_source_bitwise(wb, tag):
    g = gold("ref_dg2d.npz")
    s, steps, gcase, source = _dg2d(wb, g, tag, 1)
    with s:
        x, y = g[f"{tag}/x"], g[f"{tag}/y"]
        nodes, modes = g[f"{tag}/nodes"], g[f"{tag}/modes"]
        assert np.array_equal(s.get_modes_from_nodes(nodes), modes)
        assert np.array_equal(s.get_nodes_from_modes(modes), g[f"{tag}/nodes_back"])
        d = s.compute_update(modes, x, y)
        if source == 2 and gcase == 2:      # Keplerian grad_phi: r**(3./2.) is pow() on the device
            assert rel(d, g[f"{tag}/dudt"]) <= TOL
        else:
            assert np.array_equal(d, g[f"{tag}/dudt"]), rel(d, g[f"{tag}/dudt"])
        assert s.compute_max_speed(modes[0, 0]) == tuple(g[f"{tag}/speeds"])
        assert np.array_equal(s.apply_limiter(modes), g[f"{tag}/limited"])
        assert np.array_equal(s.apply_limiter(g[f"{tag}/rough_in"]), g[f"{tag}/rough_limited"])
        if f"{tag}/nodes_evolved" in g.files:
            un, it, t, dt = s.evolve(nodes, x, y, float(g[f"{tag}/tend"]), -1)
            assert it == steps
            if source == 2 and gcase == 2:
                assert field_err(un, g[f"{tag}/nodes_evolved"]) <= TOL
            else:
                assert np.array_equal(un, g[f"{tag}/nodes_evolved"]), field_err(un, g[f"{tag}/nodes_evolved"])


@pytest.mark.parametrize("tag", [t for t in tags("ref_dg2d.npz")])
def test_dg2d_fused_kernels_match_reference_source(wb, tag):
    """arith 0: the production stage kernel (ONP / no limiter; other limiters run the reference-order kernels)."""
    g = gold("ref_dg2d.npz")
    if f"{tag}/nodes_evolved" not in g.files:
        pytest.skip("no evolve vector")
    s, steps, gcase, source = _dg2d(wb, g, tag, 0)
    with s:
        un, it, t, dt = s.evolve(g[f"{tag}/nodes"], g[f"{tag}/x"], g[f"{tag}/y"], float(g[f"{tag}/tend"]), -1)
        assert it == steps and abs(t - float(g[f"{tag}/tend"])) <= 1e-13 * t
        assert field_err(un, g[f"{tag}/nodes_evolved"]) <= TOL


@pytest.mark.parametrize("arith", [1, 0])
@pytest.mark.parametrize("tag", tags("ref_dg2d_limiters.npz"))
def test_dg2d_limiters_on_rough_data_equal_reference_source_bitwise(wb, tag, arith):
    """arith 1: the unfused reference-order limiter kernels; arith 0: the one-pass limiter kernels of the fused flow
    (k_limiter_hio_onp, ...), which run the same operations in the same order -- both bit for bit."""
    g = gold("ref_dg2d_limiters.npz")
    n, m, bc = (int(v) for v in g[f"{tag}/meta"])
    with wb.DG2D(nx=n, ny=n, mx=m, my=m, bc=bc, limiter=str(g[f"{tag}/limiter"]), flux="llf1", ninit=1, device=0, arith=arith) as s:
        v = s.apply_limiter(g[f"{tag}/in"])
    assert np.array_equal(v, g[f"{tag}/out"]), rel(v, g[f"{tag}/out"])


@pytest.mark.parametrize("arith", [1, 0])
@pytest.mark.parametrize("tag", tags("ref_dg2d_po3.npz"))
def test_dg2d_po3_limiter_equals_reference_source_bitwise(wb, tag, arith):
    """limiter_type 'PO3' (limiter_positivity_2, 2d/limiters.f90:1587-1711): the two-pass CUDA limiter in the unfused (arith 1)
    and in the fused (arith 0) flow against the vectors produced by executing the reference text -- same doubles, and the
    same NaNs where an element's mean pressure is negative (the reference's matrix decomposition takes sqrt of it)."""
    g = gold("ref_dg2d_po3.npz")
    n, m, bc = (int(v) for v in g[f"{tag}/meta"])
    with wb.DG2D(nx=n, ny=n, mx=m, my=m, bc=bc, limiter="PO3", flux="llf1", ninit=1, device=0, arith=arith) as s:
        v = s.apply_limiter(g[f"{tag}/in"])
    assert np.array_equal(v, g[f"{tag}/out"], equal_nan=True), rel(np.nan_to_num(v), np.nan_to_num(g[f"{tag}/out"]))


# ------------------------------------------------------------------------------------------------ 1D FV
@pytest.mark.parametrize("tag", tags("ref_fv1d.npz", "fvm_"))
def test_fvm1d_cuda_equals_reference_source_bitwise(wb, tag):
    g = gold("ref_fv1d.npz")
    _, nx, bc, source, ninit, iters = (int(v) for v in g[f"{tag}/meta"])
    u0 = g[f"{tag}/u0"]
    with wb.FVM1D(nx=nx, bc=bc, source=source, device=0) as s:
        assert s.compute_max_speed(u0) == float(g[f"{tag}/cmax"])
        assert np.array_equal(s.compute_update(u0), g[f"{tag}/dudt"])
        un, it, t, dt = s.evolve(u0, float(g[f"{tag}/tend"]), -1)
    assert it == iters and (t, dt) == tuple(g[f"{tag}/clock"])
    assert np.array_equal(un, g[f"{tag}/un"])


@pytest.mark.parametrize("tag", tags("ref_fv1d.npz", "b1_"))
def test_fv1d_cuda_equals_reference_source(wb, tag):
    g = gold("ref_fv1d.npz")
    _, nx, bc, neq, ninit, iters = (int(v) for v in g[f"{tag}/meta"])
    solver = str(g[f"{tag}/solver"])
    u, weq = g[f"{tag}/u"], g[f"{tag}/weq"]
    with wb.FV1D(nx=nx, bc=bc, nequilibrium=neq, solver=solver, device=0) as s:
        assert s.compute_max_speed(u) == float(g[f"{tag}/cmax"])
        assert np.array_equal(s.compute_update_fvm(u, weq), g[f"{tag}/dudt_fvm"])      # no exp/pow in the plain scheme
        dt = float(np.float32(0.8)) * (1.0 / nx) / float(g[f"{tag}/cmax"]) / 3.0
        for key, fn in (("dudt_eql", s.compute_update), ("dudt_sr", s.compute_update_sr)):
            d = fn(u, weq)
            assert np.abs(dt * (d - g[f"{tag}/{key}"])).max() / np.abs(u).max() <= TOL, key
        un, it, t, dtl = s.evolve(u, weq, float(g[f"{tag}/tend"]), -1)
    assert it == iters and abs(t - g[f"{tag}/clock"][0]) <= 1e-13 * t
    assert rel(un, g[f"{tag}/un"]) <= TOL


# ------------------------------------------------------------------------------------------------ 1D DG
@pytest.mark.parametrize("tag", tags("ref_dg1d.npz"))
def test_dg1d_cuda_equals_reference_source(wb, tag):
    g = gold("ref_dg1d.npz")
    n, nx, riemann, source, ninit, bc, use_limiter, steps = (int(v) for v in g[f"{tag}/meta"])
    integ = str(g[f"{tag}/integrator"])
    u, du, ueq, q, ui = (g[f"{tag}/{k}"] for k in ("u", "du", "ueq", "q", "ui"))
    with wb.DG1D(n=n, nx=nx, riemann=riemann, source=source, device=0, bc=bc, use_limiter=bool(use_limiter)) as s:
        xq, wq = s.quadrature()
        assert np.array_equal(xq, g[f"{tag}/quad"][0]) and np.array_equal(wq, g[f"{tag}/quad"][1])
        cmax = s.compute_max_speed(ui)
        assert cmax == float(g[f"{tag}/cmax"])
        dt = float(np.float32(0.9)) * (1.0 / nx) / cmax / (2.0 * n + 1.0)
        scale = np.abs(ueq).max()
        d = s.compute_update_exact_delta(du, ueq)
        assert np.abs(dt * (d - g[f"{tag}/dudt_delta"])).max() / scale <= TOL
        if f"{tag}/dudt_plain" in g.files:
            d = s.compute_update(u)
            assert np.abs(dt * (d - g[f"{tag}/dudt_plain"])).max() / scale <= TOL
        if f"{tag}/dudt_exact" in g.files:
            d = s.compute_update_exact(u, q)
            ref = g[f"{tag}/dudt_exact"]
            m = ~np.isnan(ref)
            assert np.abs(dt * (d[m] - ref[m])).max() / scale <= TOL
        rough = g[f"{tag}/rough"]
        for key, fn in (("lim", s.limiter), ("lim_cons", s.limiter_cons), ("lim_tdv", s.limiter_TDV)):
            if f"{tag}/{key}" in g.files:
                assert rel(fn(rough), g[f"{tag}/{key}"]) <= TOL, key
        tend = float(g[f"{tag}/tend"])
        uu = dd = None
        if integ == "RKi":
            dd, ui2, it, t, dtl = s.evolve(du, ueq, ui, tend)
        elif integ in ("RKw", "RKe"):
            uu, dd, ui2, it, t, dtl = s.evolve_w(integ, u, du, ueq, q, ui, tend)
        else:
            uu, ui2, it, t, dtl = s.evolve_rk(integ, u, du, ueq, ui, tend)
    assert it == int(g[f"{tag}/iters"]) and abs(t - g[f"{tag}/clock"][0]) <= 1e-13 * t
    if dd is not None:
        assert np.abs(dd - g[f"{tag}/du_end"]).max() / scale <= TOL
    if uu is not None:
        assert rel(uu, g[f"{tag}/u_end"]) <= TOL
    assert rel(ui2, g[f"{tag}/ureal_end"]) <= TOL


@pytest.mark.parametrize("tag", tags("ref_dg2d_error_norms.npz"))
def test_dg2d_compute_error_cuda_equals_reference_source(wb, tag):
    """wb_dg2d_compute_error vs compute_error of the interpreted reference: max errors exact, sums to 1e-13 relative (the
    per-element accumulators are the reference's; the sum over elements is a tree instead of a sequential loop)."""
    g = gold("ref_dg2d_error_norms.npz")
    n, m, ninit = (int(v) for v in g[f"{tag}/meta"])
    with wb.DG2D(nx=n, ny=n, mx=m, my=m, ninit=ninit, device=0, arith=1) as s:
        lmax, l1, l2 = s.compute_error(g[f"{tag}/u"], g[f"{tag}/u_init"])
    assert np.array_equal(lmax, g[f"{tag}/lmax"])
    assert np.abs(l1 / g[f"{tag}/l1"] - 1).max() <= 1e-13 and np.abs(l2 / g[f"{tag}/l2"] - 1).max() <= 1e-13


def test_dg2d_compute_error_large_grid_against_numpy(wb):
    n, m = 256, 3
    rng = np.random.default_rng(2)
    u = rng.standard_normal((m, m, n, n, 4)); u0 = rng.standard_normal((m, m, n, n, 4))
    with wb.DG2D(nx=n, ny=n, mx=m, my=m, device=0) as s:
        xq, wq = s.quadrature()
        lmax, l1, l2 = s.compute_error(u, u0)
    d = u - u0
    w2 = wq[:, None] * wq[None, :]                       # [qj][qi]: both directions use the same rule
    ref1 = (np.abs(d) * w2[:, :, None, None, None]).sum(axis=(0, 1, 2, 3)) * (1.0 / n) ** 2 * 0.25
    ref2 = (d * d * w2[:, :, None, None, None]).sum(axis=(0, 1, 2, 3)) * (1.0 / n) ** 2 * 0.25
    assert np.array_equal(lmax, np.abs(d).max(axis=(0, 1, 2, 3)))
    assert np.abs(l1 / ref1 - 1).max() <= 1e-12 and np.abs(l2 / ref2 - 1).max() <= 1e-12
