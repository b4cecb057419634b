#!/usr/bin/env python
"""
bench.py -- BASELINE.json's headline: "solve wall-time & J/J' matvec HBM GB/s at m=1e7 n=1024, 1/2/4/8 B200".

A step = one complete `tralcnllss` solve of the synthetic bound-constrained GLM problem (cfg3: M = 1e7 residuals,
n = 1024 parameters, box bounds, FP64) from x0 to the reference's default tolerances: the outer augmented-Lagrangian loop
runs on the host (Python standing in for Julia) and every subproblem goes through the C ABI (`bnl_solve_subproblem`) with
HOST buffers.  Jacobian rows are sharded over the N ranks (strong scaling: M is fixed); the one collective is the
exchange of per-group sums inside the library (NVLink peer-memory stores).  Iteration counts and iterates are bit-identical
for N = 1, 2, 4, 8 (fixed row-chunk geometry), so solve time scales on IDENTICAL work.

metric  solve_wall_s (lower is better)
value   CUDA-event time inside the library summed over the subproblem solves of one tralcnllss call (inputs resident in HBM),
        max over ranks
e2e     wall-clock around the public API call, host<->device copies of x, y, fixvars included, max over ranks
roofline  the streaming Jacobian kernel (mv_stream_kernel: J'(Jv), Jv, J'w -- ONE pass over J each): algorithmic bytes of one
        pass / average launch duration (CUDA events on the library's stream, over every launch of the timed region)
cpu_baseline / --impl reference
        the reference's CPU path = the literal NumPy/OpenBLAS restatement in oracle/ (Julia is not in this image), all host
        threads, on a bounded row sample (M/256 rows, same n, full solve); `value` is that solve extrapolated linearly in M
        to the full size (every J pass is O(M n); the sample's own seconds, counts and per-pass time are reported beside it).
"""
import os
import sys

# the reference arm is CPU BLAS work: give it every core even when the launcher (torchrun) exported OMP_NUM_THREADS=1.
# This must happen before NumPy / OpenBLAS are loaded.
if "--impl" in sys.argv and "reference" in sys.argv:
    _nc = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = _nc

import argparse  # noqa: E402
import json  # noqa: E402
import statistics  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CFG = {
    # BASELINE config[2]: the headline.  Breakpoint-dominated: projected CG never runs (J'J ~ c I at M/n = 1e4)
    "cfg3": dict(M=10_000_000, n=1024, model="glm", seed=3, noise=1e-3, cond_exp=0.0),
    # same family with column scaling 10^(-j/n) and a truth vector strictly inside the box: the minor iterates (projected CG,
    # src/basic_tralcnlss.jl:690-764) carry the solve
    "cfg3cg": dict(M=10_000_000, n=1024, model="glm", seed=3, noise=1e-3, cond_exp=1.0, interior_truth=True),
    "cfg2": dict(M=1_000_000, n=256, model="expsum", seed=1, noise=1e-3, cond_exp=0.0),
    # BASELINE config[1] as SURVEY 8d words it: ONE 128-term exponential sum on one time grid (kappa(J) ~ 1e16: the reference
    # algorithm exhausts max_inner_iter on it, so it is run with --max-outer / --max-inner caps)
    "cfg2dense": dict(M=1_000_000, n=256, model="expsum_dense", seed=1, noise=1e-3, cond_exp=0.0),
    # BASELINE config[4]: ill-conditioned (column scaling 10^(-6 j/n)), J = 327.7 GB: needs >= 2 GPUs, quoted at 8; the reference
    # algorithm does not reach its tolerance on it, so it is run with iteration caps (--max-outer / --max-inner)
    "cfg5": dict(M=10_000_000, n=4096, model="glm", seed=3, noise=1e-3, cond_exp=6.0),
    # BASELINE config[3]: linear equalities + nonlinear (sphere) equality + box, AL loop exercised; general projection
    "cfg4": dict(M=4_000_000, n=2048, model="glm_mixed", m_lin=64, seed=5, noise=1e-3, cond_exp=0.0),
}
METRIC = "solve_wall_s"
UNIT = "s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def interior_truth(n, seed):
    """x_true strictly inside [-1, 1]^n (no component on a bound) for the CG workload; same hash as the models."""
    from benlsip_b200.problems import _sym

    return 0.5 * _sym(seed + 2, np.zeros(1), np.arange(n))[0]


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
def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def oracle_problem(cfg, M_s):
    from oracle.models import DenseExpSumProblem, ExpSumProblem, GlmProblem

    n = cfg["n"]
    if cfg["model"] == "expsum_dense":
        return DenseExpSumProblem(M_s, n, seed=cfg["seed"])
    if cfg["model"] == "glm":
        P = GlmProblem(M_s, n, seed=cfg["seed"], noise=cfg["noise"], cond_exp=cfg["cond_exp"])
        if cfg.get("interior_truth"):
            P.x_true = interior_truth(n, cfg["seed"])
            P.y = P._gen_y()
        return P
    return ExpSumProblem(M_s, n, seed=cfg["seed"])


def oracle_solve(cfg, M_s):
    """One full solve of the row sample by the CPU restatement, with every host thread.  Returns (seconds, trace, threads)."""
    from oracle import benlsip_oracle as O

    P = oracle_problem(cfg, M_s)
    if hasattr(P, "design"):
        P.design()  # problem generation is outside the timed region on both arms
    nthreads = host_threads()
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        ctx = threadpool_limits(limits=nthreads)
    except Exception:  # pragma: no cover
        threadpool_info = None
        ctx = None
    tr = {}
    t0 = time.perf_counter()
    O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, trace=tr)
    dt = time.perf_counter() - t0
    used = nthreads
    if threadpool_info is not None:
        used = max([d.get("num_threads", 1) for d in threadpool_info()] + [1])
    if ctx is not None:
        ctx.restore_original_limits()
    return dt, tr, used


def cpu_line(cfg, M_s, dt, tr, cores):
    c = tr["counters"]
    passes = c.get("jv", 0) + c.get("jtw", 0)
    scale = cfg["M"] / M_s
    return {"value": dt * scale, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"rows 0..{M_s} of M={cfg['M']} (M/{int(round(scale))}), n={cfg['n']}, one full tralcnllss solve; "
                      f"value = sample seconds x {scale:.0f} (every J pass is O(M n))",
            "extrapolated_full_size_solve_s": dt * scale, "sample_solve_s": dt, "sample_rows": M_s,
            "j_passes": passes, "s_per_pass_sample": dt / max(passes, 1), "cpu_matvec_GBps": 8.0 * M_s * cfg["n"] * passes / dt / 1e9,
            "counts": {"outer": tr["outer_iters"], "inner": tr["inner_iters"], "minor": tr.get("minor_iters", 0),
                       "cg": tr.get("cg_iters", 0), "breakpoints": tr.get("breakpoints", 0)}}


def run_reference(args, cfg):
    """--impl reference: rank 0 only; the other ranks exit without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    M_s = max(cfg["M"] // args.cpu_sample_div, 1024)
    for _ in range(args.warmup):
        oracle_solve(cfg, M_s)
    tot = 0.0
    last = None
    for _ in range(args.steps):
        dt, tr, cores = oracle_solve(cfg, M_s)
        tot += dt
        last = (dt, tr, cores)
    dt = tot / args.steps
    cb = cpu_line(cfg, M_s, dt, last[1], last[2])
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.config, "M": cfg["M"], "n": cfg["n"], "sample_rows": M_s, "residual_family": cfg["model"],
                       "step": "one full tralcnllss solve of the row sample; value extrapolated linearly in M to the full size"},
            "cpu_baseline": cb, "counts": cb["counts"],
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def make_solver(B, cfg, args, local_rank, world, rank, M):
    from benlsip_b200.distributed import init_solver_comm, shard_rows

    n = cfg["n"]
    row0, M_loc = shard_rows(M, world, rank)
    S = B.Solver(local_rank)
    info = S.device_info()
    need = 8.0 * M_loc * n * 1.02 + 5 * 8.0 * M_loc + (1 << 30)
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
        S.use_builtin_model({"glm": B.MODEL_GLM, "expsum": B.MODEL_EXPSUM, "expsum_dense": B.MODEL_EXPSUM_DENSE}[cfg["model"]],
                            cfg["noise"], cfg["cond_exp"], cfg["seed"])
        if cfg.get("interior_truth"):
            S.model_set_truth(interior_truth(n, cfg["seed"]))
        if args.max_outer:
            solve_kw["max_outer_iter"] = args.max_outer
        if args.max_inner:
            solve_kw["max_inner_iter"] = args.max_inner
    if world > 1:
        init_solver_comm(S)
    return S, n, M_loc, solve_kw


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    import benlsip_b200 as B

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the library has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    M = args.M or cfg["M"]
    S, n, M_loc, solve_kw = make_solver(B, cfg, args, local_rank, world, rank, M)
    x0 = S.model_vectors()["x0"]
    if args.hessian == "gram":
        S.set_hessian_mode(B.HESSIAN_GRAM)
    if args.cauchy == "literal":
        S.set_cauchy_mode(B.CAUCHY_LITERAL)

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

    def maxranks(*vals):
        if world == 1:
            return vals
        tt = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tuple(float(v) for v in tt)

    def extra_solve():  # one solve in the current mode, timed end to end (max over ranks)
        barrier()
        t0 = time.perf_counter()
        xe, tre, _ = step()
        barrier()
        (te,) = maxranks(time.perf_counter() - t0)
        return xe, tre, te

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t_begin = time.perf_counter()
    dev_ms = 0.0
    mv_ms = mv_cnt = launches = jpass = 0
    h2d = d2h = 0
    tr = x = None
    for _ in range(args.steps):
        x, tr, wall = step()
        st = tr["stats"]
        dev_ms += st["solve_ms"]
        jpass += st["j_passes"]
        mv_ms += st["hess_mul_ms"] + st["vthv_ms"] + st["jtw_ms"]
        mv_cnt += st["j_passes"] - st["gram_count"]
        launches += st["kernel_launches"]
        outer = tr["outer_iters"]
        h2d += outer * (8 * n) + 8 * ((n + 63) // 64)  # x0 per subproblem + fixvars reset
        d2h += outer * (8 * n + 8) + 8 * ((n + 63) // 64)  # x, pix per subproblem + fixvars words
    barrier()
    t_wall = time.perf_counter() - t_begin
    sampler.stop_flag = True
    t_wall, dev_ms = maxranks(t_wall, dev_ms)
    st = tr["stats"]
    counts = {"outer": tr["outer_iters"], "inner": st["inner_iters"], "minor": st["minor_iters"], "cg": st["cg_iters"],
              "breakpoints": st["breakpoints"], "hess_mul": st["hess_mul"], "vthv": st["vthv"], "jtw": st["jtw"],
              "jac_eval": st["jac_eval"], "res_eval": st["res_eval"], "j_passes": st["j_passes"],
              "cauchy_loop_launches": st["cauchy_loop_launches"], "cauchy_literal_evals": st["cauchy_literal_evals"],
              "chol_rebuilds": st["chol_rebuilds"], "chol_downdates": st["chol_downdates"], "t0_reuses": st["t0_reuses"],
              "point_reuses": st["point_reuses"], "allreduces": st["allreduces"], "p2p_allreduces": st["p2p_allreduces"],
              "mu": tr["mu"]}
    phases = {"streaming_kernels_ms": st["hess_mul_ms"] + st["vthv_ms"] + st["jtw_ms"], "jacobian_generation_ms": st["jac_eval_ms"],
              "residual_eval_ms": st["res_eval_ms"], "gram_formation_ms": st["gram_ms"], "projection_factor_ms": st["chol_ms"],
              "solve_ms": st["solve_ms"]}
    phases["other_ms"] = (phases["solve_ms"] - phases["streaming_kernels_ms"] - phases["jacobian_generation_ms"] - phases["residual_eval_ms"]
                          - phases["gram_formation_ms"] - phases["projection_factor_ms"])

    # ---- extras, reported beside (never inside) the headline ----
    extras = {}
    if not args.no_extras and cfg["model"] != "glm_mixed" and args.hessian == "matrix_free" and args.cauchy == "incremental":
        # the literal Cauchy search (a Hessian apply per breakpoint, src/basic_tralcnlss.jl:633): must give the same iterate
        S.set_cauchy_mode(B.CAUCHY_LITERAL)
        xl, trl, tl = extra_solve()
        stl = trl["stats"]
        extras["literal_cauchy_mode"] = {
            "solve_wall_s": tl, "outer": trl["outer_iters"], "inner": stl["inner_iters"], "breakpoints": stl["breakpoints"],
            "cg": stl["cg_iters"], "hess_mul": stl["hess_mul"], "j_passes": stl["j_passes"],
            "x_bitwise_equal_to_default": bool(np.array_equal(xl, x)),
            "hess_mul_avg_ms": stl["hess_mul_ms"] / max(stl["hess_mul"], 1)}
        S.set_cauchy_mode(B.CAUCHY_INCREMENTAL)
        # opt-in Gram-apply mode (G = J'J on the FP64 tensor cores once per Jacobian)
        S.set_hessian_mode(B.HESSIAN_GRAM)
        xg, trg, tg = extra_solve()
        stg = trg["stats"]
        ldp = (n + 15) // 16 * 16
        ntile = (ldp + 127) // 128
        gflops = 2.0 * M_loc * (ntile * (ntile + 1) / 2) * 128 * 128
        gavg = stg["gram_ms"] / max(stg["gram_count"], 1)
        extras["gram_mode"] = {"solve_wall_s": tg, "outer": trg["outer_iters"], "inner": stg["inner_iters"], "hess_mul": stg["hess_mul"],
                               "j_passes": stg["j_passes"], "gram_count": stg["gram_count"], "gram_ms_avg": gavg,
                               "gram_tflops_per_gpu": gflops / (gavg * 1e-3) / 1e12 if gavg > 0 else None,
                               "x_rel_diff_vs_default": float(np.linalg.norm(xg - x) / np.linalg.norm(x))}
        S.set_hessian_mode(B.HESSIAN_MATRIX_FREE)
    if rank == 0:
        peak, peak_src = peaks()
        ld = (n + 15) // 16 * 16
        alg_bytes = 8.0 * M_loc * ld + 16.0 * n  # ONE pass over the local J shard
        avg_ms = mv_ms / max(mv_cnt, 1)
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        value = dev_ms * 1e-3 / args.steps
        e2e = t_wall / args.steps
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_wall / args.steps, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.config, "M": M, "n": n, "residual_family": cfg["model"], "rows_per_gpu": M_loc,
                           "parallelism": f"row-sharded x{world}",
                           "collective": ("NVLink peer-memory exchange of per-group sums" if S.comm_info()["p2p_allreduce"] else ("ncclAllGather of per-group sums" if world > 1 else "none")),
                           "l2": f"no flush needed: the J shard streamed by every pass is {8.0 * M_loc * n / 1e9:.1f} GB >> 126 MB L2",
                           "step": "one full tralcnllss solve to the reference tolerances (defaults), outer loop on the host, subproblems through the C ABI",
                           "hessian": args.hessian, "cauchy": args.cauchy, "iteration_caps": solve_kw or None,
                           "subproblem_restart": ("r, J, J'r reused when a subproblem starts at the x the previous one ended at (bit-identical)"
                                                  if os.environ.get("BNL_REUSE_POINT", "1")[:1] != "0" else "re-evaluated (BNL_REUSE_POINT=0)")},
                "counts": counts, "phases_ms_last_step": phases,
                "final": {"pix": tr["pix"], "nb_fix": int(sum(bin(int(w)).count("1") for w in tr["fixvars_words"])),
                          "x_crc": int(np.frombuffer(x.tobytes(), dtype=np.uint64).sum() & np.uint64(0xFFFFFFFFFFFF))},
                # J / J' matvec HBM GB/s (BASELINE metric, second half): bytes of J actually streamed / time
                "matvec_hbm_GBps": achieved,
                "hbm_stream_GBps_whole_solve": 8.0 * M_loc * ld * jpass / (dev_ms * 1e-3) / 1e9,
                "roofline": {"bound": "hbm", "kernel": "mv_stream_kernel (J'(Jv) fused / Jv / J'w: one pass over J per launch)",
                             "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                             "avg_launch_ms": avg_ms, "launches_timed": mv_cnt, "algorithmic_bytes_per_launch": alg_bytes,
                             "traffic": None},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps},
                "gpu_launches": int(launches), "clocks": sampler.summary()}
        if st["cg_iters"] > 0:  # projected CG: time per iteration split into streaming and exposed latency
            hm = st["hess_mul_ms"] / max(st["hess_mul"], 1)
            line["cg"] = {"iters": st["cg_iters"], "minor_iters": st["minor_iters"], "stream_us_per_iter": 1e3 * hm,
                          # everything that is not a streaming / generator kernel (reduction tree, fused O(n) CG kernel, host
                          # decision, launch gaps), spread over the CG + minor iterations: an upper bound of the exposed latency
                          "exposed_latency_us_per_iter_upper_bound": 1e3 * phases["other_ms"] / max(st["cg_iters"] + st["minor_iters"], 1),
                          "note": "per CG iteration: one fused J'(Jv) pass + reduction tree + one fused O(n) CG kernel + one host decision"}
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if world == 1 and M == CFG["cfg3"]["M"] and n == 1024 and os.path.exists(tp):
            tj = json.load(open(tp))  # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel
            line["roofline"]["traffic"] = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
            line["roofline"]["traffic_source"] = tj["source"]
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline and cfg["model"] != "glm_mixed":
            cb = cpu_baseline(args, cfg, B, local_rank)
            if "literal_cauchy_mode" in extras:
                # the reference's own work at FULL size: J passes of the literal algorithm (counted on the GPU run of the same
                # algorithm) x the CPU's measured seconds per pass x the row ratio
                ref_passes = trl["stats"]["jv"] + trl["stats"]["jtw"]
                cb["reference_j_passes_full_size"] = ref_passes
                cb["extrapolated_full_size_solve_s_same_work"] = ref_passes * cb["s_per_pass_sample"] * (cfg["M"] / cb["sample_rows"])
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    barrier()  # no rank frees its peer-mapped mailbox while another one could still be using it
    S.close()
    if world > 1:
        dist.destroy_process_group()


def prefix_agreement(tr_cpu, inner_gpu):
    """Where the CPU port's own trajectory stops being numerically meaningful (tests/parity.py: the first inner iteration whose
    accept / trust-region / termination decision sits inside the rounding noise), and whether the GPU trajectory equals it exactly
    up to there -- the comparison the parity tests make, reported beside `counts_equal` so that a difference in the noise-driven
    tail is not mistaken for a difference in the algorithm."""
    from tests.parity import first_fragile
    inner_cpu = tr_cpu["inner"]
    F = first_fragile(dict(inner=inner_cpu))
    npre = len(inner_cpu) if F is None else F
    keys = ("k", "nb_fix", "bp_cum", "cg_cum")
    ok = len(inner_gpu) >= npre and all(
        tuple(a[k] for k in keys) == tuple(b[k] for k in keys) and abs(a["mx"] - b["mx"]) <= 1e-10 * abs(b["mx"])
        for a, b in zip(inner_gpu[:npre], inner_cpu[:npre]))
    return {"cpu_trace_noise_driven_from_inner_record": F, "records_compared_exactly": npre, "exact_prefix_equal": bool(ok),
            "criterion": "tests/parity.py first_fragile: |ared - eta*pred| <= 64 eps |mx|, pix within 1e-4 of its tolerance, or Delta <= 4 sqrt(eps)"}


def cpu_baseline(args, cfg, B, local_rank):
    """Oracle ('port') timed on this host's cores on a bounded row sample of the same workload (~10-30 s), extrapolated to the
    full size -- plus the SAME sample problem solved on the GPU: identical inputs, counts compared."""
    M_s = max(cfg["M"] // args.cpu_sample_div, 1024)
    dt, tr, cores = oracle_solve(cfg, M_s)
    cb = cpu_line(cfg, M_s, dt, tr, cores)
    try:
        T = B.Solver(local_rank)
        T.set_problem(M_s, cfg["n"])
        T.use_builtin_model({"glm": B.MODEL_GLM, "expsum": B.MODEL_EXPSUM, "expsum_dense": B.MODEL_EXPSUM_DENSE}[cfg["model"]],
                            cfg["noise"], cfg["cond_exp"], cfg["seed"])
        if cfg.get("interior_truth"):
            T.model_set_truth(interior_truth(cfg["n"], cfg["seed"]))
        x0 = T.model_vectors()["x0"]
        B.tralcnllss(x0, None, None, None, None, None, None, None, None, solver=T)  # warm-up
        T.reset_stats()
        trg = {}
        t0 = time.perf_counter()
        xg, _ = B.tralcnllss(x0, None, None, None, None, None, None, None, None, solver=T, trace=trg)
        tg = time.perf_counter() - t0
        sg = trg["stats"]
        gc = {"outer": trg["outer_iters"], "inner": sg["inner_iters"], "minor": sg["minor_iters"], "cg": sg["cg_iters"],
              "breakpoints": sg["breakpoints"]}
        cb["same_problem"] = {"rows": M_s, "gpu_solve_wall_s": tg, "cpu_solve_wall_s": dt, "gpu_counts": gc,
                              "counts_equal": gc == cb["counts"],
                              "x_rel_diff": float(np.linalg.norm(xg - tr["x"]) / np.linalg.norm(tr["x"]))}
        try:
            cb["same_problem"].update(prefix_agreement(tr, trg["inner"]))
        except Exception as e:  # pragma: no cover
            cb["same_problem"]["prefix_check_error"] = str(e)
        T.close()
    except Exception as e:  # pragma: no cover
        cb["same_problem"] = {"error": str(e)}
    return cb


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
    ap.add_argument("--cauchy", default="incremental", choices=["incremental", "literal"],
                    help="incremental = device-side guarded breakpoint loop (default; bit-identical iterates); literal = a Hessian apply per breakpoint")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra literal-Cauchy and Gram-mode solves reported beside the headline")
    args = ap.parse_args()
    cfg = dict(CFG[args.config])
    if args.M:
        cfg["M"] = args.M
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
