"""cfg5 (BASELINE config[4]): ill-conditioned GLM, M = 1e7, n = 4096 (J = 327.7 GB: needs >= 2 GPUs, run at 8),
stressing the J'J Gram formation on the FP64 tensor cores + its n^2 all-reduce + Gram-applies.
torchrun --nproc-per-node 8 tools/gram_bench.py [--M 10000000 --n 4096]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import benlsip_b200 as B
from benlsip_b200.distributed import init_solver_comm, shard_rows

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=10_000_000)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--cond-exp", type=float, default=6.0)
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
M, n = args.M, args.n
row0, m_loc = shard_rows(M, world, rank)
S = B.Solver(local)
S.set_problem(m_loc, n, M_total=M, row0=row0)
S.use_builtin_model(B.MODEL_GLM, 1e-3, args.cond_exp, 3)
if world > 1:
    init_solver_comm(S)
x0 = S.model_vectors()["x0"]
S.eval_jacobian(x0 + 0.05)
ms_kernel, flops = S.time_kernel(5, 3)            # Gram kernel alone (CUDA events, per GPU)
ms_jtjv, nbytes = S.time_kernel(0, 5)             # fused matrix-free apply for comparison
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
_, ms_ev = S.gram(want_matrix=False)              # kernel + NCCL all-reduce of n^2 doubles
torch.cuda.synchronize()
t_gram_total = time.perf_counter() - t0
S.set_hessian_mode(B.HESSIAN_GRAM)
S.reset_stats()
v = np.cos(0.1 * np.arange(n))
t0 = time.perf_counter()
for _ in range(20):
    hv = S.hess_mul(v)
t_apply = (time.perf_counter() - t0) / 20
st = S.stats()
S.set_hessian_mode(B.HESSIAN_MATRIX_FREE)
hv_mf = S.hess_mul(v)
# same-run FP64 tensor ceiling: cuBLAS DGEMM A'A (n x K x n, K rows of an FP64 matrix resident in HBM), best of 3, CUDA events
K = min(m_loc, 2_000_000_000 // (8 * n) * 4)  # <= 8 GB operand
Ad = torch.randn(K, n, dtype=torch.float64, device="cuda")
best = float("inf")
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Gd = Ad.t() @ Ad
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
dgemm_tflops = 2.0 * K * n * n / (best * 1e-3) / 1e12
del Ad, Gd
vals = torch.tensor([ms_kernel, t_gram_total, ms_jtjv], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": "cfg5 gram", "M": M, "n": n, "n_gpus": world, "rows_per_gpu": m_loc, "cond_exp": args.cond_exp,
                      "gram_kernel_ms": float(vals[0]), "gram_tflops_issued_per_gpu": flops / float(vals[0]) / 1e9,
                      "gram_tflops_2Mn2_equiv_per_gpu": 2.0 * m_loc * n * n / float(vals[0]) / 1e9,
                      "cublas_dgemm_tflops_same_run": dgemm_tflops, "cublas_dgemm_shape": [n, K, n],
                      "gram_issued_over_dgemm": flops / float(vals[0]) / 1e9 / dgemm_tflops,
                      "gram_2Mn2_equiv_over_dgemm": 2.0 * m_loc * n * n / float(vals[0]) / 1e9 / dgemm_tflops,
                      "gram_tflops_issued_total": world * flops / float(vals[0]) / 1e9,
                      "gram_plus_allreduce_wall_ms": 1e3 * float(vals[1]), "allreduce_bytes": 8 * ((n + 15) // 16 * 16) ** 2,
                      "gram_apply_device_ms": st["hess_mul_ms"] / max(st["hess_mul"], 1), "gram_apply_abi_wall_ms": 1e3 * t_apply,
                      "matrix_free_apply_ms": float(vals[2]), "matrix_free_GBps_per_gpu": nbytes / float(vals[2]) / 1e6,
                      "gram_vs_matrix_free_rel": float(np.linalg.norm(hv - hv_mf) / np.linalg.norm(hv_mf))}), flush=True)
S.close()
if world > 1:
    dist.destroy_process_group()
