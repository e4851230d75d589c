"""Host-side plumbing for the one-process-per-GPU slab decomposition (torch.distributed is only the
bootstrap: the per-stage ghost-row exchange and the CFL all-reduce run inside libwbeuler over NCCL).

Partition: 1-D slabs along y, rank r owns global rows [ny*r/R, ny*(r+1)/R) -- the same integer
arithmetic as wb_fv2d_create, so host arrays and device slabs always agree.
"""
import os

import numpy as np


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def slab_rows(ny, rank, nranks):
    """(j0, nrows) of rank's slab; identical to wb_fv2d_create's split."""
    j0 = ny * rank // nranks
    j1 = ny * (rank + 1) // nranks
    return j0, j1 - j0


def scatter_rows(global_array, rank, nranks):
    """Rows of a global (ny, nx, nvar) array owned by `rank` (a C-contiguous copy)."""
    j0, n = slab_rows(global_array.shape[0], rank, nranks)
    return np.ascontiguousarray(global_array[j0:j0 + n])


def gather_rows(local_array, ny, group=None):
    """All-gather the slabs into the global array (on every rank)."""
    import torch
    import torch.distributed as dist
    nranks = dist.get_world_size(group)
    parts = [None] * nranks
    dist.all_gather_object(parts, local_array, group=group)
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == ny
    return out


def broadcast_bytes(payload, src=0, group=None):
    """Ship a bytes object (the 128-byte NCCL unique id) from src to every rank."""
    import torch.distributed as dist
    box = [payload if dist.get_rank(group) == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def max_over_ranks(value, device=None, group=None):
    """Max of a python float over ranks (timings are always reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def make_slab_solver(cls, nranks, rank, local_rank=None, **kw):
    """Create a solver handle for this rank's slab and wire its NCCL communicator."""
    import torch.distributed as dist
    from . import nccl_get_unique_id
    dev = local_rank if local_rank is not None else rank
    s = cls(rank=rank, nranks=nranks, device=dev, **kw)
    if nranks > 1:
        uid = nccl_get_unique_id() if rank == 0 else None
        uid = broadcast_bytes(uid, 0)
        s.comm_init(uid)
        dist.barrier()
    return s
