"""Device-side rate of the 2D DG RK step (development aid): device-initialised pulse, state resident, CUDA events.
usage: dg2d_rate.py n [m] [steps] [limiter] [ninit] [source] [grad_phi_case] [bc]      env: WB_DG2D_TMA / WB_DG2D_ROWS / WB_DG2D_SCHED"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fvm-source-wb_b200"))
import torch
import wbeuler

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = int(sys.argv[2]) if len(sys.argv) > 2 else 3
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
lim = sys.argv[4] if len(sys.argv) > 4 else "ONP"
ninit = int(sys.argv[5]) if len(sys.argv) > 5 else 1
source = int(sys.argv[6]) if len(sys.argv) > 6 else 1
gpc = int(sys.argv[7]) if len(sys.argv) > 7 else 1
bc = int(sys.argv[8]) if len(sys.argv) > 8 else 1
torch.cuda.init()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    s = wbeuler.DG2D(nx=n, ny=n, mx=m, my=m, flux="llf1", limiter=lim, solver="RK4", ninit=ninit, bc=bc, source=source, grad_phi_case=gpc, device=0)
    s.set_stream(st.cuda_stream)
    s.init_device(ninit)
    s.step_async(2); s.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 0.0
    for rep in range(3):
        e0.record(st); s.step_async(steps); e1.record(st); e1.synchronize()
        ms = e0.elapsed_time(e1)
        best = max(best, n * n * 5 * steps / (ms * 1e-3))
    env = {k: v for k, v in os.environ.items() if k.startswith("WB_DG2D")}
    print(f"DG n={n} m={m} lim={lim} ninit={ninit} source={source} bc={bc} kernel={s.stage_kernel()} env={env} best {best/1e9:.3f}e9 elem-stage/s = "
          f"{best*4*m*m*8*16/5/1e9:.0f} GB/s alg = {best*4*m*m*8*16/5/1e9/6455.3:.3f} of 6455 GB/s; sim {s.sync()}")
    s.close()
