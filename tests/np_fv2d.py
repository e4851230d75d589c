"""Independent numpy restatement of benchmark_2d.f90::compute_update_exact (:465-618), written from
the Fortran text (not from oracle/fv2d.c) to cross-check the C oracle.  exp goes through math.exp
(the C library's exp, the same one the C oracle links) so that the comparison can be bit for bit.
Arrays are Fortran-shaped (nvar, nx, ny) here, on purpose, to mirror the reference's indexing."""
import math

import numpy as np

_exp = np.frompyfunc(math.exp, 1, 1)


def exp(a):
    return _exp(a).astype(np.float64)


F32 = lambda v: np.float64(np.float32(v))  # noqa: E731


def get_equilibrium_solution(neq, x, y):
    w = np.zeros((4,) + x.shape)
    if neq == 1:
        w[0] = exp(-(x + y)); w[3] = exp(-(x + y))
    elif neq in (2, 3):
        rho_0 = F32(1.21); p_0 = np.float64(1.0); g = np.float64(1.0)
        w[0] = rho_0 * exp(-(rho_0 * g / p_0) * (x + y))
        w[3] = p_0 * exp(-(rho_0 * g / p_0) * (x + y))
    return w


def compute_primitive(u, gamma):
    w = np.empty_like(u)
    w[0] = u[0]
    w[1] = u[1] / w[0]
    w[2] = u[2] / w[0]
    w[3] = (gamma - F32(1.0)) * (u[3] - 0.5 * w[0] * (w[1] ** 2 + w[2] ** 2))
    return w


def compute_conservative(w, gamma):
    u = np.empty_like(w)
    u[0] = w[0]
    u[1] = w[0] * w[1]
    u[2] = w[0] * w[2]
    u[3] = w[3] / (gamma - F32(1.)) + 0.5 * (w[0] * (w[1] ** 2 + w[2] ** 2))
    return u


def compute_flux(u, gamma):
    w = compute_primitive(u, gamma)
    f = np.empty(u.shape + (2,))
    f[0, ..., 0] = w[1] * u[0]
    f[1, ..., 0] = w[1] * u[1] + w[3]
    f[2, ..., 0] = w[0] * w[1] * w[2]
    f[3, ..., 0] = w[1] * u[3] + w[1] * w[3]
    f[0, ..., 1] = u[0] * w[2]
    f[1, ..., 1] = u[1] * w[2]
    f[2, ..., 1] = u[2] * w[2] + w[3]
    f[3, ..., 1] = w[2] * u[3] + w[2] * w[3]
    return f


def compute_speed(u, gamma):
    w = compute_primitive(u, gamma)
    cs = np.sqrt(gamma * np.maximum(w[3], 1e-10) / np.maximum(w[0], 1e-10))
    return np.sqrt(w[1] ** 2 + w[2] ** 2) + cs


def compute_llflux(ul, ur, fl, fr, gamma):
    cmax = np.maximum(compute_speed(ul, gamma), compute_speed(ur, gamma))
    return 0.5 * (fr + fl) + 0.5 * cmax * (ul - ur)


def get_source(w):
    s = np.zeros_like(w)
    s[1] = -w[0] * 1.0
    s[2] = -w[0] * 1.0
    s[3] = -w[0] * (w[1] * 1.0 + w[2] * 1.0)
    return s


def compute_update_exact(u, w_eq, nx, ny, neq, gamma, boxlen_x=1.0, boxlen_y=1.0):
    """u, w_eq: (4, nx, ny) Fortran-shaped."""
    with np.errstate(all="ignore"):
        dx = boxlen_x / np.float64(nx); dy = boxlen_y / np.float64(ny)
        odx = 1 / dx; ody = 1 / dy
        u_eq = compute_conservative(w_eq, gamma)
        delta_u = u - u_eq
        ii = np.arange(1, nx + 2)[:, None] * np.ones((1, ny + 1), dtype=np.int64)
        jj = np.ones((nx + 1, 1), dtype=np.int64) * np.arange(1, ny + 2)[None, :]
        x_faces = (ii - 1).astype(np.float64) * dx
        y_faces = (jj - 1).astype(np.float64) * dx      # sic: dx  (:513)
        x = (ii.astype(np.float32) - np.float32(0.5)).astype(np.float64) * dx
        y = (jj.astype(np.float32) - np.float32(0.5)).astype(np.float64) * dy
        w_x_faces = get_equilibrium_solution(neq, x_faces, y)
        w_y_faces = get_equilibrium_solution(neq, x, y_faces)
        u_x_faces = compute_conservative(w_x_faces, gamma)
        u_y_faces = compute_conservative(w_y_faces, gamma)
        u_left = u_x_faces[:, 0:nx, 0:ny] + delta_u
        u_right = u_x_faces[:, 1:nx + 1, 0:ny] + delta_u
        u_top = u_y_faces[:, 0:nx, 1:ny + 1] + delta_u
        u_bottom = u_y_faces[:, 0:nx, 0:ny] + delta_u
        flux_left = compute_flux(u_left, gamma); flux_right = compute_flux(u_right, gamma)
        flux_top = compute_flux(u_top, gamma); flux_bottom = compute_flux(u_bottom, gamma)
        # x sweep, faces 1..nx+1 (0-based 0..nx): left cell index clamp / right cell index clamp
        il = np.clip(np.arange(0, nx + 1) - 1, 0, nx - 1); ir = np.clip(np.arange(0, nx + 1), 0, nx - 1)
        F = compute_llflux(u_right[:, il, :], u_left[:, ir, :], flux_right[..., 0][:, il, :], flux_left[..., 0][:, ir, :], gamma)
        jl = np.clip(np.arange(0, ny + 1) - 1, 0, ny - 1); jr = np.clip(np.arange(0, ny + 1), 0, ny - 1)
        G = compute_llflux(u_top[:, :, jl], u_bottom[:, :, jr], flux_top[..., 1][:, :, jl], flux_bottom[..., 1][:, :, jr], gamma)
        s_eq = get_source(w_eq)
        s = get_source(compute_primitive(u, gamma))
        F_eq = compute_flux(u_x_faces, gamma)[..., 0]
        G_eq = compute_flux(u_y_faces, gamma)[..., 1]
        dudt = (-(F[:, 1:nx + 1, :] - F[:, 0:nx, :]) * odx
                - (G[:, :, 1:ny + 1] - G[:, :, 0:ny]) * ody
                + s - s_eq
                + (F_eq[:, 1:nx + 1, 0:ny] - F_eq[:, 0:nx, 0:ny]) * odx
                + (G_eq[:, 0:nx, 1:ny + 1] - G_eq[:, 0:nx, 0:ny]) * ody)
        dudt[:, 0, :] = 0.; dudt[:, nx - 1, :] = 0.; dudt[:, :, 0] = 0.; dudt[:, :, ny - 1] = 0.
    return dudt
