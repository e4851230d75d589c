"""world_size-2 gloo tests (CPU) of the host-side plumbing of the slab decomposition."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ny, q):
    sys.path.insert(0, os.path.join(ROOT, "fvm-source-wb_b200"))
    sys.path.insert(0, ROOT)
    from wbeuler import dist as wd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the 128-byte id travels unchanged from rank 0
        payload = bytes(range(128)) if rank == 0 else None
        got = wd.broadcast_bytes(payload, 0)
        assert got == bytes(range(128))
        # 2. slabs tile the grid; scatter -> (per-slab work with ghost rows from the neighbour) -> gather
        rng = np.random.default_rng(7)
        g = rng.standard_normal((ny, 5, 4))
        mine = wd.scatter_rows(g, rank, world)
        j0, n = wd.slab_rows(ny, rank, world)
        assert mine.shape[0] == n and np.array_equal(mine, g[j0:j0 + n])
        # ghost-row exchange pattern of the stage loop: send first/last own row to the neighbours
        lo, hi = rank - 1, rank + 1
        ghost_lo = torch.zeros(5, 4, dtype=torch.float64); ghost_hi = torch.zeros(5, 4, dtype=torch.float64)
        reqs = []
        if lo >= 0:
            reqs += [dist.isend(torch.from_numpy(mine[0].copy()), lo), dist.irecv(ghost_lo, lo)]
        if hi < world:
            reqs += [dist.isend(torch.from_numpy(mine[-1].copy()), hi), dist.irecv(ghost_hi, hi)]
        for r in reqs:
            r.wait()
        if lo >= 0:
            assert np.array_equal(ghost_lo.numpy(), g[j0 - 1])
        if hi < world:
            assert np.array_equal(ghost_hi.numpy(), g[j0 + n])
        # a 3-point vertical stencil evaluated per slab with ghosts equals the global evaluation
        ext = np.concatenate([ghost_lo.numpy()[None] if lo >= 0 else mine[:1], mine, ghost_hi.numpy()[None] if hi < world else mine[-1:]])
        loc = ext[2:] - 2 * ext[1:-1] + ext[:-2]
        gext = np.concatenate([g[:1], g, g[-1:]])
        glob = gext[2:] - 2 * gext[1:-1] + gext[:-2]
        out = wd.gather_rows(loc, ny)
        assert np.array_equal(out, glob)
        # 3. timings are reported as the max over ranks
        assert wd.max_over_ranks(1.0 + rank) == float(world)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ny", [16, 37])
def test_slab_plumbing_world2(ny):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + ny
    ps = [ctx.Process(target=_worker, args=(r, 2, port, ny, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_slab_rows_match_library_split():
    sys.path.insert(0, os.path.join(ROOT, "fvm-source-wb_b200"))
    from wbeuler import dist as wd
    for ny in (8, 37, 4096, 16384):
        for R in (1, 2, 3, 4, 8):
            rows = [wd.slab_rows(ny, r, R) for r in range(R)]
            assert rows[0][0] == 0 and sum(n for _, n in rows) == ny
            for (a, n), (b, _) in zip(rows, rows[1:]):
                assert a + n == b
