"""Row sharding of the Jacobian across the GPUs of one box: one process per GPU, rows
[row0, row0 + M_local) per rank, every O(n) quantity replicated, one exchange of the n+1 per-group sums per Hessian
apply (NVLink peer-memory stores issued inside the library).  torch.distributed is only the plumbing that carries the ncclUniqueId."""
from __future__ import annotations

import os


def shard_rows(M_total: int, nranks: int, rank: int):
    """Contiguous row range (row0, M_local) of `rank`: whole groups of the library's fixed row geometry (`bnl_shard_rows`;
    nranks in {1, 2, 4, 8}), balanced to within a few rows."""
    from . import shard_rows as _sr

    return _sr(M_total, nranks, rank)


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_solver_comm(solver, backend_group=None):
    """Creates the library's NCCL communicator: rank 0 makes the unique id, torch.distributed broadcasts it."""
    import torch
    import torch.distributed as dist

    from . import Solver

    world = dist.get_world_size()
    rank = dist.get_rank()
    if world == 1:
        solver.comm_init(1, 0, None)
        return
    if rank == 0:
        uid = Solver.comm_unique_id()
    else:
        uid = bytes(128)
    obj = [uid]
    dist.broadcast_object_list(obj, src=0, group=backend_group)
    solver.comm_init(world, rank, obj[0])
