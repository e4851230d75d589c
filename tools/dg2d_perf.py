"""Quick device-side timing of the DG 2D RK step (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fvm-source-wb_b200"))
import numpy as np
import torch
import wbeuler

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = int(sys.argv[2]) if len(sys.argv) > 2 else 3
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
lim = sys.argv[4] if len(sys.argv) > 4 else "ONP"
arith = int(sys.argv[5]) if len(sys.argv) > 5 else 0
torch.cuda.init()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    s = wbeuler.DG2D(nx=n, ny=n, mx=m, my=m, flux="llf1", limiter=lim, solver="RK4", ninit=1, device=0, arith=arith)
    s.set_stream(st.cuda_stream)
    xq, _ = s.quadrature()
    dx = 1.0 / n
    xc = (np.arange(1, n + 1) - 0.5) * dx
    x = np.empty((m, m, n, n)); y = np.empty((m, m, n, n))
    for a in range(m):
        x[:, a] = (xc + dx / 2 * xq[a])[None, None, :]
        y[a, :] = (xc + dx / 2 * xq[a])[None, :, None]
    w = np.zeros((m, m, n, n, 4)); w[..., 0] = np.exp(-((x - .5) ** 2 + (y - .5) ** 2) * 10); w[..., 1] = w[..., 0]; w[..., 2] = w[..., 0]
    w[..., 3] = w[..., 0].min() / 0.4 + w[..., 0]
    s.upload(w, x, y)
    s.step_async(2); s.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record(st); s.step_async(steps); e1.record(st); e1.synchronize()
        ms = e0.elapsed_time(e1)
        rate = n * n * 5 * steps / (ms * 1e-3)
        print(f"DG n={n} m={m} lim={lim} arith={arith} steps={steps} {ms:.2f} ms  {ms/steps:.2f} ms/step  {rate/1e6:.1f} Melem-stage/s  alg {rate*4*m*m*8*16/5/1e9:.0f} GB/s")
    print(s.sync())
    s.close()
