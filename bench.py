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


def ncu_dg_profile():
    """The committed ncu capture of the DG stage kernel (profiles/dg2d_traffic.json): dram bytes per element-stage and the
    FP64-pipe utilisation BASELINE.md section 3 asks for; {} if absent."""
    p = os.path.join(ROOT, "profiles", "dg2d_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


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
    """The reference's own algorithm on the host cores: the C oracle port of benchmark_2d.f90 (the Fortran cannot be
    compiled in this image), all host threads.  N = 1: the workload's own 4096^2 grid (same size as our arm; ~2 s per
    RK2 step on 16 cores).  N > 1: rank 0 alone, a bounded 1024^2 sample of the 16384^2 workload per step (the full grid
    would take minutes per step) -- said so in `config` and `cpu_baseline.sample`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import wb_oracle as o
    wn = args.grid or (4096 if args.gpus == 1 else 16384)
    n = args.ref_grid or (wn if args.gpus == 1 else 1024)
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
    u = o.fv2d_evolve(p, u, weq, 1e300, max(1, min(args.warmup, 2)))[0]      # warm-up: page in, spin up the thread team
    t0 = time.perf_counter()
    u = o.fv2d_evolve(p, u, weq, 1e300, args.steps)[0]
    dt = time.perf_counter() - t0
    value = n * n * 2 * args.steps / dt
    same = (n == wn)
    sample = (f"the workload's own {n}x{n} grid, {args.steps} RK2 steps" if same else
              f"{n}x{n} grid ({n*n/wn**2:.4f} of the {wn}^2 workload's cells) per step") + ", hydrostatic atmosphere + pressure bump"
    cfg = workload_config(args.gpus, args)
    cfg["reference_arm_grid"] = [n, n]
    cfg["reference_arm_same_size"] = same
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
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


def dg2d_cpu_baseline(n=64, steps=2):
    """Serial C port of 2d/benchmark_2d_dg.f90 (oracle/dg2d.c) on a bounded sample: the same pulse on an n x n grid."""
    from oracle import wb_oracle as o
    o.set_num_threads(1)
    p = o.dg2d_params(nx=n, ny=n, mx=3, my=3, flux="llf1", limiter="ONP", solver="RK4", ninit=1, bc=1)
    x, y = o.dg2d_get_coords(p)
    u0 = o.dg2d_get_initial_conditions(p, x, y)
    t0 = time.perf_counter()
    o.dg2d_evolve(p, u0, x, y, 1e300, steps)
    dt = time.perf_counter() - t0
    o.set_num_threads(o.max_threads())
    return {"value": n * n * 5 * steps / dt, "unit": "element-stage-updates/s", "cores": 1, "kind": "port",
            "sample": f"{n}x{n} elements, order 3, {steps} SSPRK(5,4) steps of the same pulse (C port of 2d/benchmark_2d_dg.f90, "
                      f"gcc -O3 -ffp-contract=off, 1 thread; host has {os.cpu_count()} cores); counted as 5 stages per step "
                      "although the reference evaluates the RHS 6 times"}


def dg2d_e2e(args, stream, n=4096, steps=3):
    """wb_dg2d_evolve(u_nodes_host, ...) with pinned host buffers: H2D of the nodal state, modes_from_nodes, `steps` SSPRK(5,4)
    steps, nodes_from_modes, D2H -- at the largest grid whose nodal array (n^2 x 36 doubles) is reasonable to pin."""
    import numpy as np
    import torch
    import wbeuler
    import ctypes as C
    lib = wbeuler.lib()
    with wbeuler.DG2D(nx=n, ny=n, mx=3, my=3, flux="llf1", limiter="ONP", solver="RK4", ninit=1, bc=1, device=0) as s:
        s.set_stream(stream.cuda_stream)
        s.init_device(1)
        u_h = torch.empty(s.shape, dtype=torch.float64).pin_memory()
        u_np = u_h.numpy()
        u0 = s.download()
        it = C.c_int(); tt = C.c_double(); dd = C.c_double()

        def evolve_call(k):      # the C-ABI call itself: the Python wrapper would copy the (pinned) array first
            st = lib.wb_dg2d_evolve(s._h, wbeuler._ptr(u_np), None, None, C.c_double(1e300), C.c_int(k), C.byref(it), C.byref(tt),
                                    C.byref(dd))
            if st != 0:
                raise RuntimeError(lib.wb_last_error().decode())
        u_np[...] = u0
        evolve_call(1)           # warm the staging buffers
        u_np[...] = u0
        del u0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        evolve_call(steps)
        el = time.perf_counter() - t0
        nbytes = u_np.nbytes
    return {"value": n * n * 5 * steps / el, "unit": "element-stage-updates/s", "h2d_bytes_per_step": nbytes / steps,
            "d2h_bytes_per_step": nbytes / steps, "seconds": el, "grid": [n, n],
            "call": f"wb_dg2d_evolve(u_nodes_host, max_iter={steps}) once at {n}x{n}: H2D nodal state from pinned memory, projection, "
                    f"{steps} SSPRK(5,4) steps, reconstruction, D2H"}


def dg2d_hio_rate(stream, n=4096, steps=4):
    import torch
    import wbeuler
    with wbeuler.DG2D(nx=n, ny=n, mx=3, my=3, flux="llf1", limiter="HIO", solver="RK4", ninit=1, bc=1, device=0) as s:
        s.set_stream(stream.cuda_stream)
        s.init_device(1)
        s.step_async(2); s.sync()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream); s.step_async(steps); e1.record(stream); e1.synchronize()
        ms = e0.elapsed_time(e1)
    return {"value": n * n * 5 * steps / (ms * 1e-3), "unit": "element-stage-updates/s", "grid": [n, n], "steps": steps,
            "note": "limiter 'HIO' + 'ONP' (2d/limiters.f90:1478-1583) after every stage: k_dg_stage_split writes the un-limited stage "
                    "result to a scratch field, k_limiter_hio_onp (reference operation order, one pass) limits it into the stage output"}


def dg2d_atmosphere_rate(args, stream, n, world=1, rank=0, local_rank=0, dev=None):
    """BASELINE config 4 as its text has it -- the perturbed hydrostatic atmosphere WITH the gravity source (IC
    2d/benchmark_2d_dg.f90:145-153, source = 2, the shipped grad_phi_case = 1): the stage kernel's source-term instantiation,
    which reads the gravity field as well (separable here: from L2-resident lines; in general +144 B per element-stage)."""
    import torch
    import wbeuler
    from wbeuler import dist as wd
    s = wd.make_slab_solver(wbeuler.DG2D, world, rank, local_rank, nx=n, ny=n, mx=3, my=3, flux="llf1", limiter="ONP", solver="RK4",
                            ninit=2, bc=1, source=2, grad_phi_case=1)
    try:
        s.set_stream(stream.cuda_stream)
        s.init_device(2)
        s.step_async(3); s.sync()
        steps = max(2, min(args.steps, 5))
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
        kern = s.stage_kernel()
    finally:
        s.close()
    peak, src = measured_peak_gbs()
    per_launch = ms * 1e-3 / (5 * steps)
    alg = 921.6 * n * n / world
    return {"value": n * n * 5 * steps / (ms * 1e-3), "unit": "element-stage-updates/s", "ms_per_step": ms / steps, "steps": steps,
            "config": {"workload": f"2D modal DG, {n}x{n} elements, order 3, SSPRK(5,4), llf1, ONP, hydrostatic atmosphere + pressure pulse "
                                   "(ninit=2, eta=0.1) with the gravity source (source=2, grad_phi_case=1, bc=1), device-initialised",
                       "grid": [n, n]},
            "roofline": {"bound": "hbm", "achieved": alg / per_launch / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / per_launch / 1e9 / peak,
                         "algorithmic_bytes_per_launch": alg, "peak_source": src,
                         "kernel": f"k_dg_stage_{kern}<3, SRC> (the same stage kernel, source-term instantiation)",
                         "note": "921.6 B per element-stage is SURVEY 8d's figure (modes only).  The gravity field of the shipped "
                                 "grad_phi_case 1 on the tensor-product grid is separable (checked bit for bit at upload): the kernel reads "
                                 "it from one row of gx and one column of gy that stay in L2; a general field streams 144 B per "
                                 "element-stage more"},
            "sim": {"iters": it, "t": t, "dt": dt}}


def dg2d_section(args, stream, world=1, rank=0, local_rank=0, dev=None):
    """BASELINE config 4 (2D modal DG order 3, SSPRK(5,4), LLF, 'ONP' limiter): element-stage updates/s with the state
    resident in HBM; 921.6 algorithmic bytes per element-stage.  Reported as an extra object of the same JSON
    line (with its own roofline / cpu_baseline / e2e); the headline metric stays the FV one.  Two workloads: the periodic
    Gaussian pulse without a source term (`value`: the stage kernel proper, comparable with round 1) and `atmosphere`: the
    perturbed hydrostatic atmosphere with the gravity source, as config 4's text has it."""
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
            s.step_async(3); s.sync()                      # W = 3 warm-up steps (15 stage launches, 0.25 s at 8192^2)
            steps = max(2, min(args.steps, 5))             # K <= 5 timed steps = 25 stage launches of 17 ms each
            sampler = ClockSampler(local_rank)
            if rank == 0:
                sampler.start()
                time.sleep(0.2)
            l0 = wbeuler.kernel_launch_count()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            if world > 1:
                import torch.distributed as dist
                dist.barrier()
            torch.cuda.synchronize()
            w0 = time.time()
            e0.record(stream); s.step_async(steps); e1.record(stream); e1.synchronize()
            w1 = time.time()
            ms = e0.elapsed_time(e1)
            if world > 1:
                ms = wd.max_over_ranks(ms, device=dev)
            clocks = sampler.stop(w0, w1) if rank == 0 else None
            it, t, dt = s.sync()
            peak, src = measured_peak_gbs()
            rate = n * n * 5 * steps / (ms * 1e-3)
            stage_launches = 5 * steps
            achieved = 921.6 * n * n / world / (ms * 1e-3 / stage_launches) / 1e9      # per GPU
            prof = ncu_dg_profile()
            kern = s.stage_kernel()
            xkind = s.exchange_kind() if world > 1 else None
            out.update({"value": rate, "ms_per_step": ms / steps, "steps": steps, "gpu_launches": wbeuler.kernel_launch_count() - l0,
                        "config": {"workload": f"2D modal DG, {n}x{n} elements, mx=my=3 (36 dof/element), SSPRK(5,4), llf1, ONP limiter, "
                                               "periodic Gaussian pulse (ninit=1), device-initialised", "grid": [n, n],
                                   "parallelism": f"y-slabs x{world} (ring), ghost rows: {xkind}" if world > 1 else "single GPU"},
                        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                     "traffic": (prof.get("dram_bytes_per_element_stage") or 0) * n * n / world or None,
                                     "kernel": f"k_dg_stage_{kern}<3> (fused update + RK combination + ONP: element split over four threads, "
                                               "every face once, rows staged by TMA; 5 launches/step)",
                                     "algorithmic_bytes_per_launch": 921.6 * n * n / world, "peak_source": src,
                                     "fp64_pipe_frac": prof.get("fp64_pipe_active_frac"),
                                     "fp64_pipe_source": prof.get("fp64_pipe_source"),
                                     "note": "the launch time includes the 4 small max-speed reduction kernels of each step; the kernel "
                                             "sits above the FP64 ridge, fp64_pipe_frac (ncu, committed capture) is the other roofline"},
                        "clocks": clocks, "sim": {"iters": it, "t": t, "dt": dt}})
            s.close()
            s = None
            try:      # config 4 with its gravity source (every rank takes part)
                out["atmosphere"] = dg2d_atmosphere_rate(args, stream, n, world, rank, local_rank, dev)
            except Exception as e:
                out["atmosphere"] = {"skipped": str(e)}
            if world == 1 and rank == 0:
                try:      # the neighbour-reading 'HIO' limiter in the fused flow (stage kernel -> scratch -> one-pass limiter kernel)
                    out["dg2d_hio"] = dg2d_hio_rate(stream)
                except Exception as e:
                    out["dg2d_hio"] = {"skipped": str(e)}
                if not args.no_cpu:
                    out["cpu_baseline"] = dg2d_cpu_baseline()
                if not args.no_e2e:
                    try:
                        out["e2e"] = dg2d_e2e(args, stream)
                    except Exception as e:
                        out["e2e"] = {"skipped": str(e)}
            return out
        except Exception as e:  # e.g. out of memory at 8192^2: fall back to the next size
            out.setdefault("skipped", []).append(f"{n}: {e}")
            if s is not None:
                s.close()
    return out


def slab_parity(world, rank, local_rank):
    """N > 1 only, before anything is timed: a small slab-decomposed run of both 2D paths must reproduce the single-GPU run
    BIT FOR BIT (every rank computes the single-GPU reference on its own device; fields are gathered over the ranks)."""
    import numpy as np
    import torch.distributed as dist
    import wbeuler
    from wbeuler import dist as wd
    res = {}
    nx, ny, steps = 512, 384, 4
    with wbeuler.FV2D(nx, ny, device=local_rank) as one:
        u, weq = one.get_initial_conditions(3)
        ref, it1, t1, dt1 = one.evolve(u, weq, 1.0, steps)
    s = wd.make_slab_solver(wbeuler.FV2D, world, rank, local_rank, nx=nx, ny=ny)
    res["fv_ghost_exchange"] = s.exchange_kind()
    got, it, t, dt = s.evolve(wd.scatter_rows(u, rank, world), wd.scatter_rows(weq, rank, world), 1.0, steps)
    full = wd.gather_rows(got, ny)
    s.close()
    ok = bool(np.array_equal(full, ref) and (it, t, dt) == (it1, t1, dt1))
    res["fv_bitwise"] = ok
    res["fv_case"] = f"FV {nx}x{ny}, {steps} RK2 steps, ninit=3: fields and (iters, t, dt) of the {world}-slab run == single-GPU run"
    n, dsteps = 64, 2
    kw = dict(nx=n, ny=n, mx=3, my=3, flux="llf1", limiter="ONP", solver="RK4", ninit=1, bc=1)
    with wbeuler.DG2D(device=local_rank, **kw) as one:
        one.init_device(1); one.step_async(dsteps); r1 = one.sync(); refm = one.download_modes()
    s = wd.make_slab_solver(wbeuler.DG2D, world, rank, local_rank, **kw)
    res["dg_ghost_exchange"] = s.exchange_kind()
    s.init_device(1); s.step_async(dsteps); r2 = s.sync()
    parts = [None] * world
    dist.all_gather_object(parts, s.download_modes())
    s.close()
    okd = bool(np.array_equal(np.concatenate(parts, axis=2), refm) and r1 == r2)
    res["dg_bitwise"] = okd
    res["dg_case"] = f"DG {n}x{n} order 3, {dsteps} SSPRK(5,4) steps, ninit=1: modes and (iters, t, dt) of the {world}-slab run == single-GPU run"
    flags = [None] * world
    dist.all_gather_object(flags, (ok, okd))
    res["fv_bitwise"] = all(f[0] for f in flags)
    res["dg_bitwise"] = all(f[1] for f in flags)
    return res


def fv2d_big_n1(args, stream, n=16384, steps=5):
    """N = 1 only: the 16384^2 grid of BASELINE config 5 on ONE GPU (30 GB resident), so that the 2/4/8-GPU lines have a
    same-grid denominator.  Also written to a scratch file the N > 1 runs of the same box read back."""
    import torch
    import wbeuler
    with wbeuler.FV2D(n, n, device=0) as s:
        s.set_stream(stream.cuda_stream)
        s.init_device(3)
        s.step_async(3); s.sync()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream); s.step_async(steps); e1.record(stream); e1.synchronize()
        ms = e0.elapsed_time(e1)
    peak, src = measured_peak_gbs()
    value = n * n * 2 * steps / (ms * 1e-3)
    achieved = ALG_BYTES_PER_CELL_STAGE * n * n / (ms * 1e-3 / (2 * steps)) / 1e9
    out = {"value": value, "unit": UNIT, "grid": [n, n], "steps": steps, "ms_per_step": ms / steps,
           "roofline_frac": achieved / peak, "note": "same grid as the N > 1 lines (config 5) on one GPU: denominator of efficiency_same_grid"}
    try:
        json.dump(out, open(N1_BIG_FILE, "w"))
    except Exception:
        pass
    return out


N1_BIG_FILE = os.path.join(os.environ.get("TMPDIR", "/tmp"), "wbeuler_fv2d_16384_n1.json")


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
    parity = slab_parity(world, rank, local_rank) if (world > 1 and not args.no_parity) else None

    solver = wd.make_slab_solver(wbeuler.FV2D, world, rank, local_rank, nx=n, ny=n)
    solver.set_stream(stream.cuda_stream)
    solver.init_device(3)
    cells_global = n * n
    if world > 1:      # how the per-stage ghost rows travel: "p2p" (stored into peer memory by the stage kernel) or "nccl"
        cfg["ghost_exchange"] = solver.exchange_kind()

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
                             "kernel": "k_stage_tma<1|2> (one fused, TMA-fed RK-stage kernel per launch, 2 per step; state resident in delta form)",
                             "algorithmic_bytes_per_launch": ALG_BYTES_PER_CELL_STAGE * cells_local_max,
                             "avg_launch_us": avg_launch_s * 1e6, "peak_source": peak_src,
                             "note": "per GPU; achieved = 80 B x cells of the largest slab / mean stage-kernel time "
                                     "(CUDA events over the timed region / launches)"},
                "sim": {"iters": iters, "t": t_sim, "dt": dt_sim, "cmax": cmax}}
        if ngpu == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        if parity is not None:
            line["slab_parity"] = parity
        if ngpu > 1:
            # same-grid efficiency against the single-GPU run of the SAME 16384^2 grid, when this box has produced one
            # (bench.py --gpus 1 writes it; the driver runs N = 1, 2, 4, 8 back to back on one box)
            try:
                n1 = json.load(open(N1_BIG_FILE))
                if n1.get("grid") == [n, n]:
                    line["efficiency_same_grid"] = value / (ngpu * n1["value"])
                    line["n1_same_grid"] = n1
            except Exception:
                line["efficiency_same_grid"] = None
            line["scaling_note"] = (f"strong scaling of the {n}^2 grid over N > 1 GPUs; the N = 1 line of this bench is BASELINE config 3 "
                                    "(4096^2), a different grid -- compare N > 1 values with `fv2d_16384_n1` of the N = 1 line "
                                    "(or `efficiency_same_grid` here), not with its headline value")
    solver.close()
    if ngpu == 1 and rank == 0 and not args.no_big:
        try:
            line["fv2d_16384_n1"] = fv2d_big_n1(args, stream)
        except Exception as e:
            line["fv2d_16384_n1"] = {"skipped": str(e)}
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
    ap.add_argument("--ref-grid", type=int, default=0, help="grid edge of the reference arm (default: the workload's 4096 at N=1, a 1024 sample at N>1)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-dg", action="store_true", help="skip the extra 2D DG (config 4) measurement at N=1")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the in-process slab == single-GPU bitwise check")
    ap.add_argument("--no-big", action="store_true", help="N=1: skip the 16384^2 single-GPU run (same-grid denominator of N>1)")
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
