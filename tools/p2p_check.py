"""Multi-GPU check (run under torchrun): the fused NVLink peer-memory all-reduce against NCCL and against a single-GPU
solve of the same (small) problem.  python -m torch.distributed.run --nproc-per-node N tools/p2p_check.py"""
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
M, n = 200_003, 1024


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
    x0 = S.model_vectors()["x0"]
    x = x0 + 0.1 * np.sin(np.arange(n))
    S.eval_jacobian(x)
    v = np.cos(0.3 * np.arange(n))
    hv = S.hess_mul(v)
    q = S.vthv(v)
    _, ss = S.residuals(x, False)
    tr = {}
    xs, _ = B.tralcnllss(x0, None, None, None, None, None, None, None, None, solver=S, trace=tr)
    res[p2p] = (hv, q, ss, xs, tr["outer_iters"], tr["stats"]["inner_iters"], tr["stats"]["p2p_allreduces"], tr["stats"]["allreduces"])
    # every rank must hold bit-identical replicated results
    t = torch.from_numpy(np.concatenate([hv, [q, ss], xs])).cuda()
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    for g in gathered:
        assert torch.equal(g, gathered[0]), "ranks disagree"
    S.close()
a, b = res[True], res[False]
rel = lambda u, w: float(np.linalg.norm(np.asarray(u) - np.asarray(w)) / np.linalg.norm(np.asarray(w)))
assert rel(a[0], b[0]) < 1e-14 and abs(a[1] - b[1]) < 1e-14 * abs(b[1]) and abs(a[2] - b[2]) < 1e-14 * abs(b[2])
# different summation orders (rank-order sum vs NCCL's tree) may flicker late inner iterations; the final iterate is then
# trajectory-dependent at the ~1e-9..1e-7 level (termination only tests the free variables)
assert rel(a[3], b[3]) < 1e-6, rel(a[3], b[3])
assert a[6] > 0 and a[6] == a[7] and b[6] == 0
if rank == 0:
    # single-GPU reference of the same problem
    os.environ["BNL_P2P_ALLREDUCE"] = "0"
    S = B.Solver(local)
    S.set_problem(M, n)
    S.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, 3)
    x0 = S.model_vectors()["x0"]
    x = x0 + 0.1 * np.sin(np.arange(n))
    S.eval_jacobian(x)
    hv1 = S.hess_mul(np.cos(0.3 * np.arange(n)))
    assert rel(a[0], hv1) < 1e-13
    print(f"p2p_check ok: world={world} p2p_allreduces={a[6]} outer/inner p2p={a[4]}/{a[5]} nccl={b[4]}/{b[5]} "
          f"hv rel p2p-vs-nccl={rel(a[0], b[0]):.2e} vs-1gpu={rel(a[0], hv1):.2e} x rel={rel(a[3], b[3]):.2e}")
    S.close()
dist.barrier()
dist.destroy_process_group()
