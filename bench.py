#!/usr/bin/env python
"""bench.py -- throughput of the explicit time-step hot path (2D well-balanced FV, SSP-RK2).

Metric (BASELINE.json): cell-updates/s per RK stage (FP64) = cells x RK stages executed / time.
One "step" = one RK2 time step (two fused stage kernels, incl. the CFL max reduction) over the grid.

  python bench.py --gpus 1 --steps K --warmup W            # 4096^2 on one B200 (BASELINE config 3)
  torchrun ... bench.py --gpus N ...                        # 16384^2 split into N y-slabs (config 5, strong scaling)
  python bench.py --impl reference ...                      # the reference's CPU algorithm (oracle port, all host threads)

The JSON line carries `value` (state resident in HBM), `e2e` (host buffers through the C-ABI call
wb_fv2d_evolve), `roofline` (HBM, algorithmic 80 B per cell-stage) and `cpu_baseline` (serial oracle port).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "fvm-source-wb_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "cell-updates/s per RK stage (FP64)"
UNIT = "cell-stage-updates/s"
ALG_BYTES_PER_CELL_STAGE = 80.0   # stage 1 = 64 B, stage 2 = 96 B per cell


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_dg_traffic_bytes_per_element_stage():
    """dram read+write per element-stage of the DG stage kernel from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "dg2d_traffic.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_element_stage"])
    except Exception:
        return None


def ncu_traffic_bytes_per_cell_stage():
    """dram read+write per cell-stage from the committed ncu capture (profiles/fv2d_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "fv2d_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_cell_stage"])
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own algorithm on the host cores: the C oracle port of benchmark_2d.f90 (the Fortran
    cannot be compiled in this image), all host threads, on a bounded sample (a 1024^2 grid of the same
    atmosphere) per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import wb_oracle as o
    n = args.ref_grid
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently
    # turn the N > 1 reference arm into a single-threaded run
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    o.set_num_threads(cores)
    p = o.fv2d_params(n, n)
    x, y = o.fv2d_get_coords(p)
    weq = o.fv2d_get_equilibrium_solution(p, x, y)
    u = o.fv2d_get_initial_conditions(p, 3, x, y)
    u = o.fv2d_evolve(p, u, weq, 1e300, max(args.warmup, 1))[0]
    t0 = time.perf_counter()
    u = o.fv2d_evolve(p, u, weq, 1e300, args.steps)[0]
    dt = time.perf_counter() - t0
    value = n * n * 2 * args.steps / dt
    wn = args.grid or (4096 if args.gpus == 1 else 16384)
    sample = f"{n}x{n} grid ({n*n/wn**2:.4f} of the {wn}^2 workload's cells) per step, hydrostatic atmosphere + pressure bump"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus, args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(ngpu, args):
    n = args.grid or (4096 if ngpu == 1 else 16384)
    return {"workload": f"2D well-balanced FV Euler+gravity (benchmark_2d.f90 evolve: SSP-RK2, LLF, equilibrium subtraction), "
                        f"{n}x{n} cells, isothermal hydrostatic atmosphere + 1e-5 pressure bump (ninit=3, nequilibrium=2)",
            "grid": [n, n], "parallelism": f"y-slabs x{ngpu}" if ngpu > 1 else "single GPU",
            "l2_policy": "inputs larger than L2 (state arrays 537 MB+ each vs 126 MB L2)",
            "rk_stages_per_step": 2}


def cpu_baseline(sample_n=1024, steps=3):
    """Serial oracle port (the reference is serial Fortran) on a bounded sample."""
    from oracle import wb_oracle as o
    o.set_num_threads(1)
    p = o.fv2d_params(sample_n, sample_n)
    x, y = o.fv2d_get_coords(p)
    weq = o.fv2d_get_equilibrium_solution(p, x, y)
    u = o.fv2d_get_initial_conditions(p, 3, x, y)
    t0 = time.perf_counter()
    o.fv2d_evolve(p, u, weq, 1e300, steps)
    dt = time.perf_counter() - t0
    o.set_num_threads(o.max_threads())
    return {"value": sample_n * sample_n * 2 * steps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{sample_n}x{sample_n} grid, {steps} RK2 steps of the same atmosphere (C port of benchmark_2d.f90, "
                      f"gcc -O3 -ffp-contract=off, 1 thread; host has {os.cpu_count()} cores)"}


def dg2d_section(args, stream, world=1, rank=0, local_rank=0, dev=None):
    """BASELINE config 4 (2D modal DG order 3, SSPRK(5,4), LLF, 'ONP' limiter, periodic pulse): element-stage updates/s
    with the state resident in HBM; 921.6 algorithmic bytes per element-stage.  Reported as an extra object of
    the same JSON line; the headline metric stays the FV one."""
    import torch
    import wbeuler
    from wbeuler import dist as wd
    out = {"metric": "element-stage updates/s (2D DG order 3, SSPRK(5,4) = 5 stages/step)", "unit": "element-stage-updates/s"}
    for n in ((args.dg_grid, 4096, 2048) if world == 1 else (args.dg_grid,)):
        s = None
        try:
            s = wd.make_slab_solver(wbeuler.DG2D, world, rank, local_rank, nx=n, ny=n, mx=3, my=3, flux="llf1", limiter="ONP",
                                    solver="RK4", ninit=1, bc=1)
            s.set_stream(stream.cuda_stream)
            s.init_device(1)
            s.step_async(2); s.sync()
            steps = max(2, min(args.steps, 5))
            l0 = wbeuler.kernel_launch_count()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            if world > 1:
                import torch.distributed as dist
                dist.barrier()
            torch.cuda.synchronize()
            e0.record(stream); s.step_async(steps); e1.record(stream); e1.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                ms = wd.max_over_ranks(ms, device=dev)
            it, t, dt = s.sync()
            peak, src = measured_peak_gbs()
            rate = n * n * 5 * steps / (ms * 1e-3)
            stage_launches = 5 * steps
            achieved = 921.6 * n * n / world / (ms * 1e-3 / stage_launches) / 1e9      # per GPU
            out.update({"value": rate, "ms_per_step": ms / steps, "steps": steps, "gpu_launches": wbeuler.kernel_launch_count() - l0,
                        "config": {"workload": f"2D modal DG, {n}x{n} elements, mx=my=3 (36 dof/element), SSPRK(5,4), llf1, ONP limiter, "
                                               "periodic Gaussian pulse (ninit=1), device-initialised", "grid": [n, n],
                                   "parallelism": f"y-slabs x{world} (ring)" if world > 1 else "single GPU"},
                        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                     "traffic": (ncu_dg_traffic_bytes_per_element_stage() or 0) * n * n / world or None,
                                     "kernel": "k_dg_stage_tma<3> (fused update + RK combination + ONP, rows staged by TMA, 5 launches/step)",
                                     "algorithmic_bytes_per_launch": 921.6 * n * n / world, "peak_source": src,
                                     "note": "the launch time includes the 4 small max-speed reduction kernels of each step"},
                        "sim": {"iters": it, "t": t, "dt": dt}})
            s.close()
            return out
        except Exception as e:  # e.g. out of memory at 8192^2: fall back to the next size
            out.setdefault("skipped", []).append(f"{n}: {e}")
            if s is not None:
                s.close()
    return out


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import wbeuler
    from wbeuler import dist as wd

    rank, world, local_rank = wd.env_rank_world()
    ngpu = args.gpus
    if world != ngpu and world != 1:
        raise SystemExit(f"--gpus {ngpu} but WORLD_SIZE={world}")
    if ngpu > 1 and world == 1:
        raise SystemExit("N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload_config(ngpu, args)
    n = cfg["grid"][0]
    stream = torch.cuda.current_stream()

    solver = wd.make_slab_solver(wbeuler.FV2D, world, rank, local_rank, nx=n, ny=n)
    solver.set_stream(stream.cuda_stream)
    solver.init_device(3)
    cells_global = n * n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident throughput: W warm-up steps, then exactly K timed steps, CUDA events on the launch stream
    solver.step_async(args.warmup)
    solver.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = wbeuler.kernel_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record(stream)
    solver.step_async(args.steps)
    e1.record(stream)
    barrier()
    w1 = time.time()
    launches = wbeuler.kernel_launch_count() - l0
    ms = e0.elapsed_time(e1)
    if world > 1:
        ms = wd.max_over_ranks(ms, device=dev)
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    iters, t_sim, dt_sim, cmax = solver.sync()
    value = cells_global * 2 * args.steps / (ms * 1e-3)

    # ---- e2e: the reference-facing call evolve(u,u_eq) with HOST (pinned) buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        shp = solver.local_shape
        u_h = torch.empty(shp, dtype=torch.float64).pin_memory()
        w_h = torch.empty(shp, dtype=torch.float64).pin_memory()
        u_np, w_np = u_h.numpy(), w_h.numpy()
        ui, wi = solver.get_initial_conditions(3)
        u_np[...] = ui; w_np[...] = wi
        del ui, wi
        import ctypes as C
        lib = wbeuler.lib()
        it = C.c_int(); tt = C.c_double(); dd = C.c_double()

        def evolve_call(k):
            st = lib.wb_fv2d_evolve(solver._h, wbeuler._ptr(u_np), wbeuler._ptr(w_np), C.c_double(1e300), C.c_int(k),
                                    C.byref(it), C.byref(tt), C.byref(dd))
            if st != 0:
                raise RuntimeError(lib.wb_last_error().decode())
        evolve_call(1)  # warm the staging buffers
        u_np[...] = solver.get_initial_conditions(3)[0]
        barrier()
        t0 = time.perf_counter()
        evolve_call(args.steps)
        barrier()
        el = time.perf_counter() - t0
        if world > 1:
            el = wd.max_over_ranks(el, device=dev)
        nbytes = u_np.nbytes
        e2e = {"value": cells_global * 2 * args.steps / el, "unit": UNIT,
               "h2d_bytes_per_step": 2 * nbytes * world / args.steps, "d2h_bytes_per_step": nbytes * world / args.steps,
               "call": f"wb_fv2d_evolve(u_host, w_eq_host, max_iter={args.steps}) once: H2D u+w_eq from pinned memory, "
                       f"{args.steps} RK2 steps, D2H u; bytes are totals/steps over all ranks", "seconds": el}
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        n_launch_stage = 2 * args.steps
        cells_local_max = max(wd.slab_rows(n, r, world)[1] for r in range(world)) * n
        avg_launch_s = ms * 1e-3 / n_launch_stage
        achieved = ALG_BYTES_PER_CELL_STAGE * cells_local_max / avg_launch_s / 1e9
        tr = ncu_traffic_bytes_per_cell_stage()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpu, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks, "gpu_launches": int(launches),
                "e2e": e2e,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": (tr * cells_local_max if tr else None),
                             "kernel": "k_stage_tma<1|2> (one fused, TMA-fed RK-stage kernel per launch, 2 per step)",
                             "algorithmic_bytes_per_launch": ALG_BYTES_PER_CELL_STAGE * cells_local_max,
                             "avg_launch_us": avg_launch_s * 1e6, "peak_source": peak_src,
                             "note": "per GPU; achieved = 80 B x cells of the largest slab / mean stage-kernel time "
                                     "(CUDA events over the timed region / launches)"},
                "sim": {"iters": iters, "t": t_sim, "dt": dt_sim, "cmax": cmax}}
        if ngpu == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
    solver.close()
    if not args.no_dg:      # BASELINE config 4 rides along as an extra object (every rank takes part in slab mode)
        dg = dg2d_section(args, stream, world, rank, local_rank, dev)
        if rank == 0:
            line["dg2d"] = dg
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=0, help="override the grid edge (default 4096 at N=1, 16384 at N>1)")
    ap.add_argument("--ref-grid", type=int, default=1024, help="grid edge of the reference arm's bounded sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-dg", action="store_true", help="skip the extra 2D DG (config 4) measurement at N=1")
    ap.add_argument("--dg-grid", type=int, default=8192)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
