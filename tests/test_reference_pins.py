"""Pins the C oracle (oracle/*.c) to the reference itself.

tests/golden/ref_*.npz were produced by executing the reference's own .f90 source text (unmodified, read from
/root/reference) with the Fortran-90 interpreter in oracle/f90interp.py -- see tests/golden/make_ref_golden.py.
Here the C restatement must reproduce those outputs on the same inputs.  The bar is BITWISE equality (`==` on
every double): the interpreter implements gfortran's arithmetic without FMA contraction and calls the same glibc
`exp`/`pow` the oracle does.  /root/reference is NOT needed to run these tests; when it is present one small case is
also re-executed live, so the committed vectors cannot drift from the script that made them.
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WB_REFERENCE", "/root/reference")


def gold(name):
    return np.load(os.path.join(HERE, "golden", name))


def tags(name):
    return sorted({k.split("/")[0] for k in gold(name).files})


def same(a, b):
    """bitwise equality of two double arrays (NaN == NaN, +0 == -0 is NOT accepted unless both are zero-valued)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


def maxdiff(a, b):
    return float(np.nanmax(np.abs(np.asarray(a) - np.asarray(b)))) if np.asarray(a).size else 0.0


# ------------------------------------------------------------------------------------------------ 2D FV
@pytest.mark.parametrize("tag", tags("ref_fv2d.npz"))
def test_fv2d_oracle_equals_reference_source(oracle, tag):
    """benchmark_2d.f90: get_coords :25-43, get_initial_conditions :45-113, get_equilibrium_solution :174-218,
    compute_primitive :145-157, compute_conservative :159-171, compute_max_speed :264-279,
    compute_update_exact :465-618, compute_update :370-463, evolve :221-260."""
    g = gold("ref_fv2d.npz")
    nx, ny, ninit, neq, steps = (int(v) for v in g[f"{tag}/meta"])
    o = oracle
    p = o.fv2d_params(nx, ny, neq)
    x, y = o.fv2d_get_coords(p)
    assert same(x, g[f"{tag}/x"]) and same(y, g[f"{tag}/y"])
    assert same(o.fv2d_get_equilibrium_solution(p, x, y), g[f"{tag}/weq"])
    assert same(o.fv2d_get_initial_conditions(p, ninit, x, y), g[f"{tag}/u_ic"])
    u, weq = g[f"{tag}/u"], g[f"{tag}/weq"]
    assert same(o.fv2d_compute_primitive(p, u), g[f"{tag}/w"])
    assert same(o.fv2d_compute_conservative(p, g[f"{tag}/w"]), g[f"{tag}/u_back"])
    assert o.fv2d_compute_max_speed(p, u) == float(g[f"{tag}/cmax"])
    d = o.fv2d_compute_update_exact(p, u, weq)
    assert same(d, g[f"{tag}/dudt"]), maxdiff(d, g[f"{tag}/dudt"])
    if f"{tag}/dudt_plain" in g.files:      # for nx < ny the reference itself indexes out of bounds (:418)
        d = o.fv2d_compute_update(p, u, weq)
        assert same(d, g[f"{tag}/dudt_plain"]), maxdiff(d, g[f"{tag}/dudt_plain"])
    un, it, t, dt, cm = o.fv2d_evolve(p, u, weq, float(g[f"{tag}/tend"]), -1)
    assert it == steps
    assert same(un, g[f"{tag}/u_evolved"]), maxdiff(un, g[f"{tag}/u_evolved"])


def test_fv2d_goldens_exercised_the_reference_routines():
    g = gold("ref_fv2d.npz")
    called = {c.split(":")[0] for c in g["random/calls"]}
    assert {"compute_update_exact", "compute_update", "compute_llflux", "compute_speed", "compute_flux", "get_source",
            "compute_max_speed", "evolve", "get_equilibrium_solution", "compute_conservative", "compute_primitive"} <= called


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only present in the build container")
def test_fv2d_live_interpretation_matches_committed_vectors():
    from oracle.f90interp import Interp
    g = gold("ref_fv2d.npz")
    tag = "ragged"
    nx, ny, ninit, neq, _ = (int(v) for v in g[f"{tag}/meta"])
    it = Interp().load(f"{REF}/parameters_2d.f90").load(f"{REF}/benchmark_2d.f90")
    it.override("parameters_2d", nx=nx, ny=ny, ninit=ninit, nequilibrium=neq)
    u = np.asfortranarray(g[f"{tag}/u"].T)
    weq = np.asfortranarray(g[f"{tag}/weq"].T)
    dudt = np.zeros_like(u)
    it.call("compute_update_exact", u, weq, dudt)
    assert same(dudt.T, g[f"{tag}/dudt"])


# ------------------------------------------------------------------------------------------------ 1D FV
@pytest.mark.parametrize("tag", [t for t in tags("ref_fv1d.npz") if t.startswith("fvm_")])
def test_fvm1d_oracle_equals_reference_source(oracle, tag):
    """fvm.f90: condinit :99-174, compute_update :188-251 (+ compute_source, compute_llflux, compute_flux,
    compute_speed), compute_max_speed :320-336, the RK2 main loop of program fvm :56-76."""
    g = gold("ref_fv1d.npz")
    _, nx, bc, source, ninit, iters = (int(v) for v in g[f"{tag}/meta"])
    o = oracle
    p = o.fvm1d_params(nx=nx, bc=bc, source=source)
    u0 = g[f"{tag}/u0"]
    assert same(o.fvm1d_initial_conditions(p, ninit), u0)
    assert o.fvm1d_compute_max_speed(p, u0) == float(g[f"{tag}/cmax"])
    d = o.fvm1d_compute_update(p, u0)
    assert same(d, g[f"{tag}/dudt"]), maxdiff(d, g[f"{tag}/dudt"])
    un, it, t, dt = o.fvm1d_evolve(p, u0, float(g[f"{tag}/tend"]), -1)
    assert it == iters and (t, dt) == tuple(g[f"{tag}/clock"])
    assert same(un, g[f"{tag}/un"]), maxdiff(un, g[f"{tag}/un"])


@pytest.mark.parametrize("tag", [t for t in tags("ref_fv1d.npz") if t.startswith("b1_")])
def test_fv1d_oracle_equals_reference_source(oracle, tag):
    """benchmark_1d.f90: get_x :24-38, get_initial_conditions :40-67, get_equilibrium_solution :127-155,
    compute_max_speed :157-170, compute_update ('EQL') :263-377, compute_update_fvm :454-549,
    compute_update_sr ('WB1') :553-747, evolve :200-261."""
    g = gold("ref_fv1d.npz")
    _, nx, bc, neq, ninit, iters = (int(v) for v in g[f"{tag}/meta"])
    solver = str(g[f"{tag}/solver"])
    o = oracle
    p = o.fv1d_params(nx=nx, bc=bc, nequilibrium=neq, solver=solver)
    x = o.fv1d_get_x(p)
    assert same(x, g[f"{tag}/x"])
    assert same(o.fv1d_get_equilibrium_solution(p, x), g[f"{tag}/weq"])
    assert same(o.fv1d_get_initial_conditions(p, ninit, x, float(g[f"{tag}/eta"])), g[f"{tag}/u_ic"])
    u, weq = g[f"{tag}/u"], g[f"{tag}/weq"]
    assert o.fv1d_compute_max_speed(p, u) == float(g[f"{tag}/cmax"])
    for key, fn in (("dudt_eql", o.fv1d_compute_update), ("dudt_fvm", o.fv1d_compute_update_fvm), ("dudt_sr", o.fv1d_compute_update_sr)):
        d = fn(p, u, weq)
        assert same(d, g[f"{tag}/{key}"]), (key, maxdiff(d, g[f"{tag}/{key}"]))
    un, it, t, dt = o.fv1d_evolve(p, u, weq, float(g[f"{tag}/tend"]), -1)
    assert it == iters and (t, dt) == tuple(g[f"{tag}/clock"])
    assert same(un, g[f"{tag}/un"]), maxdiff(un, g[f"{tag}/un"])


# ------------------------------------------------------------------------------------------------ 1D DG
def _nan_aware_same(got, ref):
    """Entries the reference computes from memory outside its arrays (faces 1 and nx+1 without an index clamp,
    dg_with_source.f90:1903-1914) are NaN in the interpreted run; everything else must agree bit for bit."""
    ref = np.asarray(ref)
    m = ~np.isnan(ref)
    return got.shape == ref.shape and np.array_equal(np.asarray(got)[m], ref[m]), int((~m).sum())


@pytest.mark.parametrize("tag", tags("ref_dg1d.npz"))
def test_dg1d_oracle_equals_reference_source(oracle, tag):
    """dg_with_source.f90: set-up of program dg :26-171 (projections with root legendre.f90's single-precision basis and
    quadrature), compute_max_speed :1136-1153, compute_update_exact_delta :1749-2031, compute_update :807-1028,
    compute_update_exact :1380-1744, limiter :414-519, limiter_TDV :523-608, limiter_cons :610-734,
    riemann_llf/hllc :1299-1374, and the main loop :173-336 for every integrator."""
    g = gold("ref_dg1d.npz")
    n, nx, riemann, source, ninit, bc, use_limiter, steps = (int(v) for v in g[f"{tag}/meta"])
    integ = str(g[f"{tag}/integrator"])
    o = oracle
    p = o.dg1d_params(n=n, nx=nx, riemann=riemann, source=source, ninit=ninit, pert=float(g[f"{tag}/pert"]), bc=bc,
                      use_limiter=use_limiter)
    xq, wq = o.dg1d_quadrature(p)
    assert same(xq, g[f"{tag}/quad"][0]) and same(wq, g[f"{tag}/quad"][1])
    ui, ueq, du = o.dg1d_setup(p)
    assert same(ui, g[f"{tag}/ui"]) and same(ueq, g[f"{tag}/ueq"])
    assert same(du, g[f"{tag}/du"]), maxdiff(du, g[f"{tag}/du"])
    u, q = o.dg1d_project(p, ui), o.dg1d_project(p, ueq)
    assert same(u, g[f"{tag}/u"]) and same(q, g[f"{tag}/q"])
    assert o.dg1d_compute_max_speed(p, ui) == float(g[f"{tag}/cmax"])
    d = o.dg1d_compute_update_exact_delta(p, du, ueq)
    assert same(d, g[f"{tag}/dudt_delta"]), maxdiff(d, g[f"{tag}/dudt_delta"])
    if f"{tag}/dudt_plain" in g.files:
        d = o.dg1d_compute_update(p, u)
        assert same(d, g[f"{tag}/dudt_plain"]), maxdiff(d, g[f"{tag}/dudt_plain"])
    if f"{tag}/dudt_exact" in g.files:
        d = o.dg1d_compute_update_exact(p, u, q)
        ok, nans = _nan_aware_same(d, g[f"{tag}/dudt_exact"])
        assert ok and nans <= 2 * 3 * n, (nans, maxdiff(d, g[f"{tag}/dudt_exact"]))
    rough = g[f"{tag}/rough"]
    for key, fn in (("lim", o.dg1d_limiter), ("lim_cons", o.dg1d_limiter_cons), ("lim_tdv", o.dg1d_limiter_tdv)):
        if f"{tag}/{key}" in g.files:
            v = fn(p, rough)
            assert same(v, g[f"{tag}/{key}"]), (key, maxdiff(v, g[f"{tag}/{key}"]))
    tend = float(g[f"{tag}/tend"])
    if integ == "RKi":
        dd, ui2, it, t, dt = o.dg1d_evolve_rki(p, du, ueq, ui, tend)
        uu = None
    elif integ in ("RKw", "RKe"):
        uu, dd, ui2, it, t, dt = o.dg1d_evolve_w(p, integ, u, du, ueq, q, ui, tend)
    else:
        uu, ui2, it, t, dt = o.dg1d_evolve_rk(p, integ, u, du, ueq, ui, tend)
        dd = None
    assert it == int(g[f"{tag}/iters"]) and (t, dt) == tuple(g[f"{tag}/clock"])
    if dd is not None:
        ok, nans = _nan_aware_same(dd, g[f"{tag}/du_end"])
        assert ok, maxdiff(dd, g[f"{tag}/du_end"])
    if uu is not None:
        ok, nans = _nan_aware_same(uu, g[f"{tag}/u_end"])
        assert ok, maxdiff(uu, g[f"{tag}/u_end"])
    ok, nans = _nan_aware_same(ui2, g[f"{tag}/ureal_end"])
    assert ok, maxdiff(ui2, g[f"{tag}/ureal_end"])


# ------------------------------------------------------------------------------------------------ 2D DG
def _dg2d_params(o, g, tag):
    n, m, bc, source, gcase, ninit, steps = (int(v) for v in g[f"{tag}/meta"])
    flux, lim, solver = (str(s) for s in g[f"{tag}/names"])
    box = float(g[f"{tag}/boxlen"]) if f"{tag}/boxlen" in g.files else 1.0
    return o.dg2d_params(nx=n, ny=n, mx=m, my=m, bc=bc, source=source, grad_phi_case=gcase, flux=flux, limiter=lim,
                         solver=solver, ninit=ninit, boxlen_x=box, boxlen_y=box), steps


@pytest.mark.parametrize("tag", tags("ref_dg2d.npz"))
def test_dg2d_oracle_equals_reference_source(oracle, tag):
    """2d/benchmark_2d_dg.f90: get_coords :93-120, get_initial_conditions :122-466, get_modes_from_nodes :497-542,
    get_nodes_from_modes :544-592, compute_update :1137-1479 (+ compute_flux, compute_flux_int, compute_num_flux,
    compute_llflux / compute_hllflux / compute_hllcflux, get_boundary_conditions, get_source, grad_phi,
    special_boundary_conditions), compute_max_speed :826-870, apply_limiter :1516-1555, evolve :624-775;
    2d/legendre.f90 legendre, legendre_prime, gl_quadrature, gll_quadrature; 2d/limiters.f90 compute_positivity,
    compute_set, solve_for_t, high_order_limiter, limiting, minmod2d, generalized_minmod, compute_limiter,
    limiter_low_order."""
    g = gold("ref_dg2d.npz")
    o = oracle
    p, steps = _dg2d_params(o, g, tag)
    x, y = o.dg2d_get_coords(p)
    assert same(x, g[f"{tag}/x"]) and same(y, g[f"{tag}/y"])
    nodes = o.dg2d_get_initial_conditions(p, x, y)          # all twelve cases are restated
    assert same(nodes, g[f"{tag}/nodes"]), maxdiff(nodes, g[f"{tag}/nodes"])
    nodes = g[f"{tag}/nodes"]
    modes = o.dg2d_get_modes_from_nodes(p, nodes)
    assert same(modes, g[f"{tag}/modes"])
    assert same(o.dg2d_get_nodes_from_modes(p, modes), g[f"{tag}/nodes_back"])
    d = o.dg2d_compute_update(p, modes, x, y)
    assert same(d, g[f"{tag}/dudt"]), maxdiff(d, g[f"{tag}/dudt"])
    assert o.dg2d_compute_max_speed(p, modes) == tuple(g[f"{tag}/speeds"])
    v = o.dg2d_apply_limiter(p, modes)
    assert same(v, g[f"{tag}/limited"]), maxdiff(v, g[f"{tag}/limited"])
    v = o.dg2d_apply_limiter(p, g[f"{tag}/rough_in"])
    assert same(v, g[f"{tag}/rough_limited"]), maxdiff(v, g[f"{tag}/rough_limited"])
    if f"{tag}/nodes_evolved" in g.files:
        un, it, t, dt = o.dg2d_evolve(p, nodes, x, y, float(g[f"{tag}/tend"]), -1)
        assert it == steps and t == float(g[f"{tag}/tend"])
        assert same(un, g[f"{tag}/nodes_evolved"]), maxdiff(un, g[f"{tag}/nodes_evolved"])


@pytest.mark.parametrize("tag", tags("ref_dg2d_po3.npz"))
def test_dg2d_po3_limiter_equals_reference_source(oracle, tag):
    """limiter_type 'PO3' = limiter_positivity_2 (2d/limiters.f90:1587-1711) with compute_characteristics,
    compute_cons_from_characteristics and get_matrix_decomp (2d/benchmark_2d_dg.f90:2021-2177), executed from the reference
    text on rough and on smooth modes.  Where an element's mean pressure is negative the sound speed of its matrix
    decomposition is NaN and the reference returns NaN for that element: reproduced (same doubles, same NaNs)."""
    g = gold("ref_dg2d_po3.npz")
    n, m, bc = (int(v) for v in g[f"{tag}/meta"])
    p = oracle.dg2d_params(nx=n, ny=n, mx=m, my=m, bc=bc, limiter="PO3", flux="llf1", ninit=1)
    v = oracle.dg2d_apply_limiter(p, g[f"{tag}/in"])
    assert np.array_equal(v, g[f"{tag}/out"], equal_nan=True), maxdiff(v, g[f"{tag}/out"])
    assert not np.array_equal(v, g[f"{tag}/in"], equal_nan=True)


@pytest.mark.parametrize("tag", tags("ref_dg2d_ics.npz"))
def test_dg2d_initial_conditions_equal_reference_source(oracle, tag):
    """get_initial_conditions (2d/benchmark_2d_dg.f90:122-466), cases 3..12 -- Riemann problems, isentropic vortex, the two
    rotating disks, square and 1-d pulse advection, the Gaussian with w(4) = minval(w(1)), the Keplerian disk with the
    softened potential -- executed from the reference text: the oracle's restatement gives the same doubles."""
    g = gold("ref_dg2d_ics.npz")
    ninit, n, m = (int(v) for v in g[f"{tag}/meta"])
    box = float(g[f"{tag}/boxlen"])
    p = oracle.dg2d_params(nx=n, ny=n, mx=m, my=m, ninit=ninit, boxlen_x=box, boxlen_y=box)
    x, y = oracle.dg2d_get_coords(p)
    assert same(x, g[f"{tag}/x"]) and same(y, g[f"{tag}/y"])
    nodes = oracle.dg2d_get_initial_conditions(p, x, y)
    assert same(nodes, g[f"{tag}/nodes"]), maxdiff(nodes, g[f"{tag}/nodes"])


@pytest.mark.parametrize("tag", tags("ref_dg2d_limiters.npz"))
def test_dg2d_limiters_on_rough_data_equal_reference_source(oracle, tag):
    """apply_limiter on data where every limiter acts: negative density / pressure at points of the Zhang-Shu set
    (2d/limiters.f90 compute_positivity :478-654), steep modes (high_order_limiter :1478-1583, compute_limiter :203-309,
    limiter_low_order :769-860)."""
    g = gold("ref_dg2d_limiters.npz")
    n, m, bc = (int(v) for v in g[f"{tag}/meta"])
    p = oracle.dg2d_params(nx=n, ny=n, mx=m, my=m, bc=bc, limiter=str(g[f"{tag}/limiter"]), flux="llf1", ninit=1)
    v = oracle.dg2d_apply_limiter(p, g[f"{tag}/in"])
    assert not same(g[f"{tag}/in"], g[f"{tag}/out"])
    assert same(v, g[f"{tag}/out"]), maxdiff(v, g[f"{tag}/out"])


def test_reference_test_program_test2d_as_shipped(oracle):
    """2d/test2d.f90, the reference's only test program, interpreted as shipped (nx = ny = 8, mx = my = 2): exp(-x+y) at
    the GL nodes, projection, reconstruction, then 1000 round trips through 2d/commons.f90's get_modes_from_nodes /
    get_nodes_from_modes; it prints maxval(u - nodes) and minval(u - nodes).  The oracle must land on the same nodes,
    bit for bit, after the same 1000 trips (and therefore print the same two numbers)."""
    g = gold("ref_test2d.npz")
    o = oracle
    p = o.dg2d_params(nx=8, ny=8, mx=2, my=2)
    x, y = o.dg2d_get_coords(p)
    u = g["u"]
    assert u.shape == (2, 2, 8, 8, 4)
    # test2d builds x with dx/dble(2) instead of get_coords' dx/2.0 -- the same number; u must be exp(-x+y) of those points
    ref_u = np.empty_like(u)
    import math
    for idx in np.ndindex(x.shape):
        ref_u[idx] = math.exp(-x[idx] + y[idx])
    assert same(u, ref_u)
    modes = o.dg2d_get_modes_from_nodes(p, u)
    nodes = o.dg2d_get_nodes_from_modes(p, modes)
    for _ in range(1000):
        modes = o.dg2d_get_modes_from_nodes(p, nodes)
        nodes = o.dg2d_get_nodes_from_modes(p, modes)
    assert same(nodes, g["nodes"]), maxdiff(nodes, g["nodes"])
    assert same(modes, g["modes"])
    printed = np.array([np.max(u - nodes), np.min(u - nodes)])
    assert same(printed, g["printed"]) and 0 < printed[1] < printed[0] < 2e-12
    called = {c.split(":")[0]: int(c.split(":")[1]) for c in g["calls"]}
    assert called["get_modes_from_nodes"] == 1000 and called["get_nodes_from_modes"] == 1000


@pytest.mark.parametrize("tag", tags("ref_dg2d_error_norms.npz"))
def test_dg2d_compute_error_equals_reference_source(oracle, tag):
    """compute_error (2d/benchmark_2d_dg.f90:23-89): the L1 / L2 accumulators of the interpreted reference, bit for bit
    (sequential sums in the reference's loop order), and the max errors it prints."""
    g = gold("ref_dg2d_error_norms.npz")
    n, m, ninit = (int(v) for v in g[f"{tag}/meta"])
    p = oracle.dg2d_params(nx=n, ny=n, mx=m, my=m, ninit=ninit)
    u_init = oracle.dg2d_get_initial_conditions(p, g[f"{tag}/x"], g[f"{tag}/y"])
    assert same(u_init, g[f"{tag}/u_init"])
    lmax, l1, l2 = oracle.dg2d_compute_error(p, g[f"{tag}/u"], u_init)
    assert same(l1, g[f"{tag}/l1"]) and same(l2, g[f"{tag}/l2"]) and same(lmax, g[f"{tag}/lmax"])


def test_limiter_branches_that_are_not_built_and_why():
    """apply_limiter (2d/benchmark_2d_dg.f90:1516-1555) has ten branches.  ONP, HIO, 1OR, LOW and POS are built.  What the
    interpreted reference does with the other five (tests/golden/ref_dg2d_other_limiters.json): 'ROS' indexes u_avg(.., 0),
    'KRI' passes a 4-element section where a whole 5-D array is expected, '1DL' calls a subroutine that does not exist --
    undefined behaviour or no program at all; 'COC' ends by overwriting the nodal pressure with the literal 10e-5 (an
    abandoned experiment); 'PO3' runs (characteristic-variable minmod) and is built (test_dg2d_po3_limiter_equals_reference_source)."""
    import json
    f = json.load(open(os.path.join(HERE, "golden", "ref_dg2d_other_limiters.json")))
    for k, v in f.items():
        lim = k.split("_")[0]
        if lim in ("ROS", "KRI", "1DL"):
            assert v["status"] == "error"
        else:
            assert v["status"] == "ran"
        if lim == "ROS":
            assert "u_avg" in v["message"] and "out of bounds" in v["message"]
        if lim == "KRI":
            assert "get_nodes_from_modes" in v["message"]
        if lim == "1DL":
            assert "limiter_1d" in v["message"]
        if lim == "COC":
            assert abs(v["pressure_min"] - 1e-4) < 1e-11 and abs(v["pressure_max"] - 1e-4) < 1e-11


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only present in the build container")
@pytest.mark.parametrize("src,ranges,gone,kept", [
    ("benchmark_2d.f90", ["221-260", "264-279", "465-618"], {"evolve", "compute_max_speed", "compute_update_exact"},
     {"main", "get_coords", "get_initial_conditions", "get_equilibrium_solution", "compute_flux", "output_file"}),
    ("2d/benchmark_2d_dg.f90", ["23-89", "497-592", "624-775", "826-870", "1137-1479", "1516-1555"],
     {"compute_error", "get_modes_from_nodes", "get_nodes_from_modes", "evolve", "compute_max_speed", "compute_update", "apply_limiter"},
     {"main", "get_coords", "get_initial_conditions", "get_equilibrium_solution", "output_file", "compute_num_flux"}),
    # 1D programs: fvm.f90 and dg_with_source.f90 keep their time loop in the main program -> replaced by one call
    ("fvm.f90", ["188-251", "320-336", "56-76=  call wb_fvm1d_time_loop(u,t,dt,iter)"], {"compute_update", "compute_max_speed"},
     {"fvm", "condinit", "compute_primitive", "compute_llflux", "compute_flux", "compute_speed"}),
    ("benchmark_1d.f90", ["157-170", "200-261", "263-377", "454-549", "553-747"],
     {"compute_max_speed", "evolve", "compute_update", "compute_update_fvm", "compute_update_sr"},
     {"main", "get_x", "get_initial_conditions", "get_equilibrium_solution", "output_file", "compute_llflux"}),
    ("dg_with_source.f90", ["414-519", "523-606", "610-734", "807-1028", "1136-1152", "1380-1744", "1749-2031",
                            "173-336=  call wb_dg1d_time_loop(u,delta_u,u_eq,u_eq_modes,uinit,t,dt,iter)"],
     {"limiter", "limiter_tdv", "limiter_cons", "compute_update", "compute_max_speed", "compute_update_exact",
      "compute_update_exact_delta"},
     {"dg", "condinit", "get_eq_solution", "modes_to_nodes", "nodes_to_modes", "riemann_hllc"}),
])
def test_splitter_ranges_of_the_integration_recipes_cut_whole_routines(tmp_path, src, ranges, gone, kept):
    """INTEGRATION.md / the Fortran shims tell a maintainer which line ranges tools/split_reference.py must drop.  The image
    has no Fortran compiler, so the result is parsed with the Fortran-90 front end of oracle/f90interp.py instead: the
    remainder must still be a sequence of complete program units, the replaced routines must be gone and the driver parts
    (program, initialisers, output) must still be there."""
    import subprocess
    import sys
    from oracle.f90interp import Interp
    out = tmp_path / "driver.f90"
    subprocess.check_call([sys.executable, os.path.join(os.path.dirname(HERE), "tools", "split_reference.py"),
                           os.path.join(REF, src), str(out)] + ranges)
    it = Interp().load(str(out))
    names = set(it.units)
    assert not (gone & names), gone & names
    assert kept <= names, kept - names
    text = out.read_text()
    for r in ranges:                 # a replaced time loop: the call into the shim stands where the loop was
        if "=" in r:
            call = r.split("=", 1)[1].strip()
            assert call in text
            first = int(r.split("-")[0])
            assert "do while" in open(os.path.join(REF, src)).read().splitlines()[first - 1]
