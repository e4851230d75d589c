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
