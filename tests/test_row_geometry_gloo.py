"""World-size-2 gloo test (CPU) of the N > 1 reduction logic: the fixed row-chunk geometry (csrc/rowgeom.h, `bnl_shard_rows`)
and its summation tree -- teams/chunks -> group sums -> the 8 group sums in order -- restated in NumPy.  Two ranks each reduce
the chunks of their own groups, exchange ONLY the group sums (all_gather, like the NVLink mailbox / ncclAllGather of the
library), and add the 8 vectors in group order: the result must equal the single-rank evaluation of the same tree BIT FOR BIT,
for J'(Jv), ||Jv||^2 and ||r||^2 -- the property that makes iteration counts independent of the GPU count."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KGROUPS, KCHAINS = 8, 8


def chunk_begin(M, G, c):
    n = KGROUPS * G
    base, extra = divmod(M, n)
    return c * base + min(c, extra)


def pick_G(M):
    return max(1, min(148, M // 512))


def group_sums(P, row0, M_total, groups, x, v, n):
    """Group sums of [J'(Jv), ||Jv||^2, ||r||^2] for the given groups; P holds the rows [row0, ...) of the global problem."""
    G = pick_G(M_total)
    J, r = P.jac_res(x), P.residuals(x)
    out = {}
    for g in groups:
        chains = [np.zeros(n + 2) for _ in range(KCHAINS)]
        for b in range(G):
            lo, hi = chunk_begin(M_total, G, g * G + b) - row0, chunk_begin(M_total, G, g * G + b + 1) - row0
            Jc, rc = J[lo:hi].copy(), r[lo:hi].copy()  # fresh buffers: the same bytes at the same alignment on every rank
            t = Jc @ v
            part = np.concatenate([Jc.T @ t, [t @ t], [rc @ rc]])  # one partial per chunk (a CTA's fixed pattern)
            chains[b % KCHAINS] = chains[b % KCHAINS] + part       # kChains interleaved chains over the chunks of the group
        s = chains[0]
        for k in range(1, KCHAINS):
            s = s + chains[k]
        out[g] = s
    return out


def in_order(sums):
    s = sums[0]
    for g in range(1, KGROUPS):
        s = s + sums[g]
    return s


def _worker(rank, world, port, M, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from benlsip_b200.distributed import shard_rows
    from oracle.models import GlmProblem

    row0, m_loc = shard_rows(M, world, rank)
    ng = KGROUPS // world
    assert row0 == chunk_begin(M, pick_G(M), rank * ng * pick_G(M))  # the library's shard = whole groups of the geometry
    P = GlmProblem(m_loc, n, seed=3, row0=row0)
    x, v = np.linspace(-0.5, 0.5, n), np.cos(np.arange(n))
    mine = group_sums(P, row0, M, range(rank * ng, (rank + 1) * ng), x, v, n)
    send = torch.from_numpy(np.stack([mine[g] for g in sorted(mine)]))
    got = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(got, send)  # only group sums cross ranks
    allsums = np.concatenate([t.numpy() for t in got])
    res = in_order(list(allsums))
    if rank == 0:
        q.put(res.copy())
    dist.barrier()
    dist.destroy_process_group()


def test_group_sum_exchange_is_bitwise_rank_count_invariant():
    from oracle.models import GlmProblem

    M, n, world = 20_011, 24, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    P = GlmProblem(M, n, seed=3)
    x, v = np.linspace(-0.5, 0.5, n), np.cos(np.arange(n))
    one = in_order([group_sums(P, 0, M, range(KGROUPS), x, v, n)[g] for g in range(KGROUPS)])
    assert np.array_equal(got, one)  # bit for bit
    J, r = P.jac_res(x), P.residuals(x)
    ref = np.concatenate([J.T @ (J @ v), [(J @ v) @ (J @ v)], [r @ r]])
    np.testing.assert_allclose(got, ref, rtol=1e-12)
