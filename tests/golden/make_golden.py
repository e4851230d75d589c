"""Generates the committed golden vectors under tests/golden/ from the C oracle (oracle/*.c).

The reference itself cannot be run here (no Fortran compiler in the image), so these vectors pin the
ORACLE against regressions; the oracle in turn is pinned by the analytic invariants in
tests/test_oracle_*.py and by independent numpy restatements.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wb_oracle as o  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fv2d():
    out = {}
    for tag, nx, ny, ninit, neq in (("sq_pert", 24, 24, 3, 2), ("ragged_pert", 33, 20, 3, 2),
                                    ("riemann", 40, 40, 4, 2), ("riemann_ragged", 32, 24, 4, 2), ("eq1", 20, 24, 1, 1)):
        p = o.fv2d_params(nx, ny, neq)
        x, y = o.fv2d_get_coords(p)
        weq = o.fv2d_get_equilibrium_solution(p, x, y)
        u = o.fv2d_get_initial_conditions(p, ninit, x, y)
        out[f"{tag}_meta"] = np.array([nx, ny, ninit, neq])
        out[f"{tag}_u"] = u
        out[f"{tag}_weq"] = weq
        out[f"{tag}_dudt"] = o.fv2d_compute_update_exact(p, u, weq)
        out[f"{tag}_dudt_plain"] = o.fv2d_compute_update(p, u, weq)
        un, it, t, dt, cm = o.fv2d_evolve(p, u, weq, 1.0, 3)
        out[f"{tag}_u3"] = un
        out[f"{tag}_clock"] = np.array([it, t, dt, cm])
    np.savez_compressed(os.path.join(HERE, "fv2d.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["fv2d"]
    for w in which:
        globals()[w]()
        print("wrote", w)
