#!/usr/bin/env python
"""
bench.py -- BASELINE.json's headline: "solve wall-time & J/J' matvec HBM GB/s at m=1e7 n=1024, 1/2/4/8 B200".

A step = one complete `tralcnllss` solve of the synthetic bound-constrained GLM problem (cfg3: M = 1e7 residuals,
n = 1024 parameters, box bounds, FP64) from x0: the outer augmented-Lagrangian loop runs on the host (Python
standing in for Julia) and every subproblem goes through the C ABI (`bnl_solve_subproblem`) with HOST buffers.
Jacobian rows are sharded over the N ranks (strong scaling: M is fixed), the one collective is the library's NCCL
all-reduce of n+1 doubles per Hessian apply.

metric  matvec_equiv_GBps: 8*M*n bytes per J.v or J'.w product the ALGORITHM performs (a Hessian apply = 2 products,
        exactly what the reference's two DGEMVs stream) divided by time.  Same accounting on both arms.
        NOTE the fused kernel reads J once per Hessian apply, so `value` can exceed the HBM peak; the HBM-honest
        number is `roofline` (algorithmic bytes of ONE pass / measured kernel time).
value   from CUDA-event time inside the library around each subproblem solve (inputs resident in HBM)
e2e     same work / wall-clock around the public API call, host<->device copies of x, y, fixvars included
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CFG = {"cfg3": dict(M=10_000_000, n=1024, model="glm", seed=3, noise=1e-3, cond_exp=0.0),
       "cfg2": dict(M=1_000_000, n=256, model="expsum", seed=1, noise=1e-3, cond_exp=0.0),
       # BASELINE config[3]: linear equalities + nonlinear (sphere) equality + box, AL loop exercised; general projection
       "cfg4": dict(M=4_000_000, n=2048, model="glm_mixed", m_lin=64, seed=5, noise=1e-3, cond_exp=0.0)}
METRIC = "matvec_equiv_GBps"
UNIT = "GB/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4) if s[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------
def run_reference(args, cfg):
    """The reference's CPU implementation of the path (the literal NumPy/OpenBLAS restatement in oracle/, since Julia is
    not in this image -- DESIGN.md) on a bounded row sample of the same workload, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import benlsip_oracle as O
    from oracle.models import ExpSumProblem, GlmProblem

    M_s = max(cfg["M"] // args.cpu_sample_div, 1024)
    n = cfg["n"]
    P = GlmProblem(M_s, n, seed=cfg["seed"]) if cfg["model"] == "glm" else ExpSumProblem(M_s, n, seed=cfg["seed"])
    try:
        from threadpoolctl import threadpool_info
        cores = max([d.get("num_threads", 1) for d in threadpool_info()] + [1])
    except Exception:
        cores = os.cpu_count()

    def step():
        tr = {}
        t0 = time.perf_counter()
        O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, trace=tr)
        dt = time.perf_counter() - t0
        c = tr["counters"]
        return dt, 8.0 * M_s * n * (c.get("jv", 0) + c.get("jtw", 0)), tr

    for _ in range(args.warmup):
        step()
    tot_t = tot_b = 0.0
    tr = None
    for _ in range(args.steps):
        dt, b, tr = step()
        tot_t += dt
        tot_b += b
    val = tot_b / tot_t / 1e9
    sample = f"rows 0..{M_s} of M={cfg['M']} (M/{args.cpu_sample_div}), n={n}, full tralcnllss solve per step"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.config, "M": cfg["M"], "n": n, "sample_rows": M_s, "residual_family": cfg["model"],
                       "bytes_accounting": "8*M*n per J.v or J'.w product"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "counts": {"outer": tr["outer_iters"], "inner": tr["inner_iters"], "cg": tr.get("cg_iters", 0)}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    import benlsip_b200 as B
    from benlsip_b200.distributed import init_solver_comm, shard_rows

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the library has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    M, n = cfg["M"], cfg["n"]
    if args.M:
        M = args.M
    row0, M_loc = shard_rows(M, world, rank)
    S = B.Solver(local_rank)
    info = S.device_info()
    need = 8.0 * M_loc * n * 1.02 + 3 * 8.0 * M_loc + (1 << 30)
    if need > info["free_bytes"]:
        raise SystemExit(f"J shard ({need/1e9:.1f} GB) does not fit the GPU ({info['free_bytes']/1e9:.1f} GB free)")
    solve_kw = {}
    if cfg["model"] == "glm_mixed":
        from benlsip_b200.problems import mixed_constraint_setup
        n = args.n or n
        mc = mixed_constraint_setup(n, args.m_lin or cfg["m_lin"], cfg["seed"])
        S.set_problem(M_loc, n, mc["A"], mc["xlow"], mc["xupp"], p=1, M_total=M, row0=row0)
        S.use_builtin_model(B.MODEL_GLM, cfg["noise"], cfg["cond_exp"], cfg["seed"])
        S.model_set_truth(mc["x_star"], mc["x0"])
        S.use_builtin_nlcons(B.NLCONS_SPHERE, mc["rho2"])
        solve_kw = dict(max_outer_iter=args.max_outer or 500, max_inner_iter=args.max_inner or 500)
    else:
        S.set_problem(M_loc, n, M_total=M, row0=row0)
        S.use_builtin_model(B.MODEL_GLM if cfg["model"] == "glm" else B.MODEL_EXPSUM, cfg["noise"], cfg["cond_exp"], cfg["seed"])
    if world > 1:
        init_solver_comm(S)
    x0 = S.model_vectors()["x0"]
    if args.hessian == "gram":
        S.set_hessian_mode(B.HESSIAN_GRAM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        tr = {}
        S.reset_stats()
        t0 = time.perf_counter()
        x, _ = B.tralcnllss(x0, None, None, None, None, None, None, None, None, solver=S, trace=tr, **solve_kw)  # public API, host buffers
        wall = time.perf_counter() - t0
        return x, tr, wall

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t_begin = time.perf_counter()
    dev_ms = 0.0
    prod_bytes = 0.0
    hm_ms = hm_cnt = launches = jpass = 0
    h2d = d2h = 0
    tr = x = None
    for _ in range(args.steps):
        x, tr, wall = step()
        st = tr["stats"]
        dev_ms += st["solve_ms"]
        prod_bytes += 8.0 * M * n * (st["jv"] + st["jtw"])
        jpass += st["j_passes"]
        hm_ms += st["hess_mul_ms"]
        hm_cnt += st["hess_mul"]
        launches += st["kernel_launches"]
        outer = tr["outer_iters"]
        h2d += outer * (8 * n) + 8 * ((n + 63) // 64)  # x0 per subproblem + fixvars reset
        d2h += outer * (8 * n + 8) + 8 * ((n + 63) // 64)  # x, pix per subproblem + fixvars words
    barrier()
    t_wall = time.perf_counter() - t_begin
    sampler.stop_flag = True
    # extra: one solve in the opt-in Gram-apply mode (G = J'J on the FP64 tensor cores once per Jacobian), reported beside
    # the headline, never mixed into it
    gram_extra = None
    if args.hessian == "matrix_free" and not args.no_gram_extra and cfg["model"] != "glm_mixed":
        S.set_hessian_mode(B.HESSIAN_GRAM)
        step()  # warm-up (allocations)
        barrier()
        tg0 = time.perf_counter()
        xg, trg, _ = step()
        barrier()
        tg = time.perf_counter() - tg0
        stg = trg["stats"]
        ldp = (n + 15) // 16 * 16
        ntile = (ldp + 127) // 128
        gflops = 2.0 * M_loc * (ntile * (ntile + 1) / 2) * 128 * 128
        gram_extra = {"solve_wall_s": tg, "outer": trg["outer_iters"], "inner": stg["inner_iters"], "hess_mul": stg["hess_mul"],
                      "j_passes": stg["j_passes"], "gram_count": stg["gram_count"],
                      "gram_ms_avg": stg["gram_ms"] / max(stg["gram_count"], 1),
                      "gram_tflops_per_gpu": gflops / (stg["gram_ms"] / max(stg["gram_count"], 1) * 1e-3) / 1e12 if stg["gram_ms"] > 0 else None,
                      "x_rel_diff_vs_matrix_free": float(np.linalg.norm(xg - x) / np.linalg.norm(x)),
                      "matvec_equiv_GBps": 8.0 * M * n * (stg["jv"] + stg["jtw"]) / tg / 1e9}
        S.set_hessian_mode(B.HESSIAN_MATRIX_FREE)
    # extra: one solve with the opt-in incremental Cauchy search (no Hessian apply per breakpoint)
    inc_extra = None
    if args.hessian == "matrix_free" and not args.no_gram_extra and cfg["model"] != "glm_mixed":
        S.set_cauchy_mode(B.CAUCHY_INCREMENTAL)
        step()
        barrier()
        ti0 = time.perf_counter()
        xi, tri, _ = step()
        barrier()
        ti = time.perf_counter() - ti0
        sti = tri["stats"]
        inc_extra = {"solve_wall_s": ti, "outer": tri["outer_iters"], "inner": sti["inner_iters"], "breakpoints": sti["breakpoints"],
                     "hess_mul": sti["hess_mul"], "j_passes": sti["j_passes"],
                     "x_rel_diff_vs_default": float(np.linalg.norm(xi - x) / np.linalg.norm(x)),
                     "matvec_equiv_GBps_literal_accounting": 8.0 * M * n * (st["jv"] + st["jtw"]) / ti / 1e9}
        S.set_cauchy_mode(B.CAUCHY_LITERAL)
    # max over ranks of both clocks
    if world > 1:
        tt = torch.tensor([t_wall, dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_wall, dev_ms = float(tt[0]), float(tt[1])
    if rank == 0:
        peak, peak_src = peaks()
        ld = (n + 15) // 16 * 16
        alg_bytes = 8.0 * M_loc * ld + 16.0 * n  # ONE pass over the local J shard
        avg_ms = hm_ms / max(hm_cnt, 1)
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        value = prod_bytes / (dev_ms * 1e-3) / 1e9
        e2e = prod_bytes / t_wall / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_wall / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.config, "M": M, "n": n, "residual_family": cfg["model"], "rows_per_gpu": M_loc,
                           "parallelism": f"row-sharded x{world}", "collective": ("fused NVLink peer-memory all-reduce" if S.comm_info()["p2p_allreduce"] else ("nccl" if world > 1 else "none")), "l2": f"no flush needed: the J shard streamed by every apply is {8.0 * M_loc * n / 1e9:.1f} GB >> 126 MB L2",
                           "bytes_accounting": "8*M*n per J.v or J'.w product; a fused Hessian apply = 2 products, 1 HBM pass",
                           "step": "one full tralcnllss solve to the reference tolerances (defaults)"},
                "solve_wall_s": t_wall / args.steps, "solve_device_s": dev_ms * 1e-3 / args.steps,
                # HBM-honest whole-solve figure: bytes of J actually streamed (one pass per fused apply) / device time
                "hbm_stream_GBps": 8.0 * M * ((n + 15) // 16 * 16) * jpass / (dev_ms * 1e-3) / 1e9,
                "counts": {"outer": tr["outer_iters"], "inner": tr["stats"]["inner_iters"], "minor": tr["stats"]["minor_iters"],
                           "cg": tr["stats"]["cg_iters"], "breakpoints": tr["stats"]["breakpoints"], "hess_mul": tr["stats"]["hess_mul"],
                           "vthv": tr["stats"]["vthv"], "jtw": tr["stats"]["jtw"], "jac_eval": tr["stats"]["jac_eval"],
                           "chol_rebuilds": tr["stats"]["chol_rebuilds"], "mu": tr["mu"],
                           "res_eval": tr["stats"]["res_eval"], "j_passes": tr["stats"]["j_passes"], "allreduces": tr["stats"]["allreduces"],
                           "p2p_allreduces": tr["stats"]["p2p_allreduces"]},
                "final": {"pix": tr["pix"], "nb_fix": int(sum(bin(int(w)).count("1") for w in tr["fixvars_words"]))},
                "roofline": {"bound": "hbm", "kernel": "mv_stream_kernel<JTJV> (fused J'(Jv), one pass)", "achieved": achieved,
                             "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                             "avg_launch_ms": avg_ms, "launches_timed": hm_cnt, "algorithmic_bytes_per_launch": alg_bytes,
                             "traffic": None},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps},
                "gpu_launches": int(launches), "clocks": sampler.summary()}
        line["config"]["hessian"] = args.hessian
        tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if world == 1 and M == CFG["cfg3"]["M"] and n == 1024 and args.hessian == "matrix_free" and os.path.exists(tp):
            tj = json.load(open(tp))  # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel
            line["roofline"]["traffic"] = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
            line["roofline"]["traffic_source"] = tj["source"]
        if gram_extra is not None:
            line["gram_mode"] = gram_extra
        if inc_extra is not None:
            line["incremental_cauchy_mode"] = inc_extra
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, cfg)
        print(json.dumps(line), flush=True)
    S.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, cfg):
    """Oracle ('port') timed on this host's cores on a bounded row sample of the same workload (~10-30 s)."""
    if cfg["model"] == "glm_mixed":
        return None
    from oracle import benlsip_oracle as O
    from oracle.models import ExpSumProblem, GlmProblem

    M_s = max(cfg["M"] // args.cpu_sample_div, 1024)
    n = cfg["n"]
    P = GlmProblem(M_s, n, seed=cfg["seed"]) if cfg["model"] == "glm" else ExpSumProblem(M_s, n, seed=cfg["seed"])
    try:
        from threadpoolctl import threadpool_info
        cores = max([d.get("num_threads", 1) for d in threadpool_info()] + [1])
    except Exception:
        cores = os.cpu_count()
    tr = {}
    t0 = time.perf_counter()
    O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, trace=tr)
    dt = time.perf_counter() - t0
    c = tr["counters"]
    val = 8.0 * M_s * n * (c.get("jv", 0) + c.get("jtw", 0)) / dt / 1e9
    return {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
            "sample": f"rows 0..{M_s} of M={cfg['M']} (M/{args.cpu_sample_div}), n={n}, one full tralcnllss solve",
            "counts": {"outer": tr["outer_iters"], "inner": tr["inner_iters"], "cg": tr.get("cg_iters", 0)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CFG))
    ap.add_argument("--M", type=int, default=0, help="override the row count (debug)")
    ap.add_argument("--n", type=int, default=0, help="override n (cfg4 only, debug)")
    ap.add_argument("--m-lin", type=int, default=0, dest="m_lin")
    ap.add_argument("--max-outer", type=int, default=0, dest="max_outer")
    ap.add_argument("--max-inner", type=int, default=0, dest="max_inner")
    ap.add_argument("--cpu-sample-div", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hessian", default="matrix_free", choices=["matrix_free", "gram"],
                    help="matrix_free = the reference's J'(Jv) semantics (default, parity mode); gram = opt-in DMMA Gram-apply mode")
    ap.add_argument("--no-gram-extra", action="store_true", help="skip the extra (untimed-in-headline) Gram-mode solve")
    args = ap.parse_args()
    cfg = CFG[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
