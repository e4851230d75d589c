import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fvm-source-wb_b200"))
import wbeuler as wb
g = np.load(os.path.join(ROOT, "tests/golden/ref_dg2d.npz"))
tag = "advsink_o2"
n, m, bc, source, gcase, ninit, steps = (int(v) for v in g[f"{tag}/meta"])
flux, lim, solver = (str(s) for s in g[f"{tag}/names"])
nodes, x, y = g[f"{tag}/nodes"], g[f"{tag}/x"], g[f"{tag}/y"]
for MAXIT in (-1, 2, 1):
 for src in (3, 1):
   for lm in ("ONP", "none"):
     for sv in ("SS4", "DEB", "EQL"):
         res = {}
         for arith in (0, 1):
             with wb.DG2D(nx=n, ny=n, mx=m, my=m, bc=bc, source=src, grad_phi_case=gcase, flux=flux, limiter=lm, solver=sv, ninit=ninit, device=0, arith=arith) as s:
                 res[arith] = s.evolve(nodes, x, y, float(g[f"{tag}/tend"]), MAXIT)
         a, b = res[0], res[1]
         print(f"maxit={MAXIT} source={src} limiter={lm} solver={sv}: it {a[1]} {b[1]} dt {a[3]:.6e} {b[3]:.6e} err {np.abs(a[0]-b[0]).max()/np.abs(b[0]).max():.3e}")
