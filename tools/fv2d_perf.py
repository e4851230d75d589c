"""Quick device-side timing of the fused FV2D stage kernels (development aid; bench.py is the contract)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fvm-source-wb_b200"))
import torch
import wbeuler

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
arith = int(sys.argv[3]) if len(sys.argv) > 3 else 0
torch.cuda.init()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    s = wbeuler.FV2D(n, n, arith=arith, device=0)
    s.set_stream(st.cuda_stream)
    s.init_device(3)
    s.step_async(5); s.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record(st); s.step_async(steps); e1.record(st); e1.synchronize()
        ms = e0.elapsed_time(e1); best = min(best, ms)
        rate = n * n * 2 * steps / (ms * 1e-3)
        print(f"n={n} arith={arith} steps={steps} {ms:.3f} ms  {ms/steps/2*1e3:.1f} us/stage  {rate/1e9:.2f} Gcell-stage/s  alg {rate*80/1e9:.0f} GB/s")
    print(s.sync())
    s.close()
