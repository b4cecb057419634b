"""Multi-GPU check (run under torchrun with 2, 4 or 8 ranks): the NVLink peer-memory exchange of the per-group sums against
the ncclAllGather fallback and against a SINGLE-GPU solve of the same problem -- all three must agree BIT FOR BIT (fixed
row-chunk geometry, csrc/rowgeom.h), including the whole solve trajectory (iteration counts, final iterate, active set).
python -m torch.distributed.run --nproc-per-node N tools/p2p_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import benlsip_b200 as B
from benlsip_b200.distributed import init_solver_comm, shard_rows

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
M, n = int(os.environ.get("P2P_CHECK_M", 200_003)), int(os.environ.get("P2P_CHECK_N", 1024))
x_probe = 0.1 * np.sin(np.arange(n))
v = np.cos(0.3 * np.arange(n))


def say(*a):
    if os.environ.get("P2P_CHECK_VERBOSE"):
        print(f"[rank {rank}]", *a, file=sys.stderr, flush=True)


def run(S):
    x0 = S.model_vectors()["x0"]
    S.eval_jacobian(x0 + x_probe)
    hv = S.hess_mul(v)
    q = S.vthv(v)
    _, ss = S.residuals(x0 + x_probe, False)
    g = S.gradient(x0 + x_probe)
    say("kernels ok", S.comm_info())
    S.eval_jacobian(x0)
    s1, pred1 = S.inner_step(x0, S.gradient(x0), 0.5)
    say("inner_step ok", pred1, S.stats()["breakpoints"], S.stats()["cauchy_loop_launches"])
    tr = {}
    xs, _ = B.tralcnllss(x0, None, None, None, None, None, None, None, None, solver=S, trace=tr)
    st = tr["stats"]
    counts = (tr["outer_iters"], st["inner_iters"], st["minor_iters"], st["cg_iters"], st["breakpoints"])
    return np.concatenate([hv, [q, ss], g, xs]), counts, tr["fixvars_words"], st


def make(p2p):
    os.environ["BNL_P2P_ALLREDUCE"] = "1" if p2p else "0"
    S = B.Solver(local)
    row0, m_loc = shard_rows(M, world, rank)
    S.set_problem(m_loc, n, M_total=M, row0=row0)
    S.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, 3)
    init_solver_comm(S)
    return S


res = {}
for p2p in (True, False):
    S = make(p2p)
    info = S.comm_info()
    assert info["p2p_allreduce"] == p2p, info
    res[p2p] = run(S)
    # every rank must hold bit-identical replicated results
    t = torch.from_numpy(res[p2p][0]).cuda()
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    for gth in gathered:
        assert torch.equal(gth, gathered[0]), "ranks disagree"
    S.close()
a, b = res[True], res[False]
assert np.array_equal(a[0], b[0]) and a[1] == b[1] and np.array_equal(a[2], b[2]), "peer-memory exchange vs ncclAllGather"
assert a[3]["p2p_allreduces"] > 0 and a[3]["p2p_allreduces"] == a[3]["allreduces"] and b[3]["p2p_allreduces"] == 0
if rank == 0:
    # single-GPU solve of the same problem: the trajectory must not depend on the GPU count
    S = B.Solver(local)
    S.set_problem(M, n)
    S.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, 3)
    one = run(S)
    S.close()
    assert np.array_equal(one[0], a[0]), float(np.max(np.abs(one[0] - a[0])))
    assert one[1] == a[1] and np.array_equal(one[2], a[2])
    print(f"p2p_check ok: world={world} M={M} n={n} counts(outer,inner,minor,cg,bp)={a[1]} identical at 1 and {world} GPUs, "
          f"peer-memory == allgather == single GPU bit for bit; exchanges={a[3]['p2p_allreduces']}")
dist.barrier()
dist.destroy_process_group()
