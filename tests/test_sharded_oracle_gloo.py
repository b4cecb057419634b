"""World-size-2 gloo test (CPU) of the N>1 host logic: the row sharding used by the multi-GPU path
(benlsip_b200.distributed.shard_rows + the models' row0 offset) composes -- per-rank partial J'(Jv), ||Jv||^2 and
||r||^2, all-reduced, equal the single-rank values (the only collective of the path)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, M, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from benlsip_b200.distributed import shard_rows
    from oracle.models import GlmProblem

    row0, m_loc = shard_rows(M, world, rank)
    P = GlmProblem(m_loc, n, seed=3, row0=row0)
    x = np.linspace(-0.5, 0.5, n)
    v = np.cos(np.arange(n))
    J = P.jac_res(x)
    r = P.residuals(x)
    Jv = J @ v
    buf = torch.from_numpy(np.concatenate([J.T @ Jv, [Jv @ Jv], [r @ r], J.T @ r]))
    dist.all_reduce(buf)  # the path's only collective: sum of the n+1 (+ g) partials
    if rank == 0:
        q.put(buf.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharding_allreduce_matches_single_rank():
    from oracle.models import GlmProblem

    M, n, world = 1001, 24, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    P = GlmProblem(M, n, seed=3)
    x = np.linspace(-0.5, 0.5, n)
    v = np.cos(np.arange(n))
    J, r = P.jac_res(x), P.residuals(x)
    Jv = J @ v
    ref = np.concatenate([J.T @ Jv, [Jv @ Jv], [r @ r], J.T @ r])
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
