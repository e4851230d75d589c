"""Device-side rate of the 2D DG RK step (development aid): device-initialised pulse, state resident, CUDA events.
usage: dg2d_rate.py n [m] [steps] [limiter]      env: WB_DG2D_SPLIT / WB_DG2D_TMA / WB_DG2D_MARCH / WB_DG2D_ROWS select the kernel"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fvm-source-wb_b200"))
import torch
import wbeuler

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = int(sys.argv[2]) if len(sys.argv) > 2 else 3
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
lim = sys.argv[4] if len(sys.argv) > 4 else "ONP"
torch.cuda.init()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    s = wbeuler.DG2D(nx=n, ny=n, mx=m, my=m, flux="llf1", limiter=lim, solver="RK4", ninit=1, bc=1, device=0)
    s.set_stream(st.cuda_stream)
    s.init_device(1)
    s.step_async(2); s.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 0.0
    for rep in range(3):
        e0.record(st); s.step_async(steps); e1.record(st); e1.synchronize()
        ms = e0.elapsed_time(e1)
        best = max(best, n * n * 5 * steps / (ms * 1e-3))
    env = {k: v for k, v in os.environ.items() if k.startswith("WB_DG2D")}
    print(f"DG n={n} m={m} lim={lim} kernel={s.stage_kernel()} env={env} best {best/1e9:.3f}e9 elem-stage/s = "
          f"{best*4*m*m*8*16/5/1e9:.0f} GB/s alg = {best*4*m*m*8*16/5/1e9/6455.3:.3f} of 6455 GB/s; sim {s.sync()}")
    s.close()
