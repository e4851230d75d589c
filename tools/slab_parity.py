"""Slab-decomposed evolve (one rank per GPU, NCCL ghost-row exchange) vs single-GPU evolve vs the CPU oracle.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/slab_parity.py [nx ny steps]
Exit code 0 = parity holds on every rank."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fvm-source-wb_b200"))
import numpy as np
import torch
import torch.distributed as dist
import wbeuler
from wbeuler import dist as wd

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 96
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 70
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
rank, world, local_rank = wd.env_rank_world()
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
ok = True
kinds = set()
for arith in (0, 1):
    # ghost rows by peer-memory stores of the stage kernel / by NCCL send-recv overlapped with the interior / not overlapped
    for overlap, p2p in (("1", "1"), ("1", "0"), ("0", "0")):
        os.environ["WB_FV2D_OVERLAP"] = overlap
        os.environ["WB_FV2D_P2P"] = p2p
        s = wd.make_slab_solver(wbeuler.FV2D, world, rank, local_rank, nx=nx, ny=ny, arith=arith)
        kinds.add(s.exchange_kind())
        # global fields from the library's own generator on a single-GPU handle (every rank, deterministic)
        with wbeuler.FV2D(nx, ny, arith=arith, device=local_rank) as one:
            u, weq = one.get_initial_conditions(3)
            ref_single, it1, t1, dt1 = one.evolve(u, weq, 1.0, steps)
            c_single = one.compute_max_speed(u)
        mine_u, mine_w = wd.scatter_rows(u, rank, world), wd.scatter_rows(weq, rank, world)
        got, it, t, dt = s.evolve(mine_u, mine_w, 1.0, steps)
        full = wd.gather_rows(got, ny)
        c_slab = s.compute_max_speed(mine_u)
        d = s.compute_update_exact(mine_u, mine_w)
        dfull = wd.gather_rows(d, ny)
        with wbeuler.FV2D(nx, ny, arith=arith, device=local_rank) as one:
            dref = one.compute_update_exact(u, weq)
        same = np.array_equal(full, ref_single) and it == it1 and t == t1 and c_slab == c_single and np.array_equal(dfull, dref)
        msg = (f"rank {rank}/{world} arith={arith} overlap={overlap} exchange={s.exchange_kind()}: slab == single-GPU bitwise: {same} "
               f"(iters {it}, t {t:.6e})")
        if rank == 0:
            from oracle import wb_oracle as o
            p = o.fv2d_params(nx, ny)
            oref = o.fv2d_evolve(p, u, weq, 1.0, steps)[0]
            err = np.abs(full - oref).max() / np.abs(oref).max()
            msg += f"; vs oracle rel Linf {err:.2e}"
            same = same and err <= 1e-12
        print(msg, flush=True)
        ok = ok and same
        s.close()
        dist.barrier()
if rank == 0:
    print("exchange kinds exercised:", sorted(kinds), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
