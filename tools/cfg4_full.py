"""cfg4 at BASELINE size (M = 4e6, n = 2048, m_lin = 64 linear equalities + sphere constraint + box; J = 65.5 GB) on ONE GPU,
end to end through the library's outer loop (bnl_tralcnllss), default modes (matrix-free Hessian, Gram-guarded Cauchy search).
The benlsip.out-format log is flushed every inner iteration, so a cut-off run still shows how far it got.
    python tools/cfg4_full.py [M n m_lin max_outer max_inner]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import benlsip_b200 as B
from benlsip_b200.problems import mixed_constraint_setup

M, n, m_lin = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4_000_000, 2048, 64)
max_outer = int(sys.argv[4]) if len(sys.argv) > 4 else 500
max_inner = int(sys.argv[5]) if len(sys.argv) > 5 else 500
out = os.environ.get("CFG4_OUT", "gpurun_out/r2_cfg4_full")
S = B.Solver(0)
mc = mixed_constraint_setup(n, m_lin, 5)
S.set_problem(M, n, mc["A"], mc["xlow"], mc["xupp"], p=1)
S.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, 5)
S.model_set_truth(mc["x_star"], mc["x0"])
S.use_builtin_nlcons(B.NLCONS_SPHERE, mc["rho2"])
S.set_params(max_inner_iter=max_inner)
S.reset_stats()
t0 = time.perf_counter()
x, y, mu, pix = S.tralcnllss_native(mc["x0"], log_path=out + ".log", max_outer_iter=max_outer)
wall = time.perf_counter() - t0
st = S.stats()
c, _ = S.nlcons(x)
res = {"workload": "cfg4 full size", "M": M, "n": n, "m_lin": m_lin, "p": 1, "solve_wall_s": wall, "final_mu": mu, "final_pix": pix,
       "nl_feasibility": float(abs(c[0])), "lin_feasibility": float(np.max(np.abs(mc["A"] @ x - mc["b"]))),
       "x_dist_to_truth_rel": float(np.linalg.norm(x - mc["x_star"]) / np.linalg.norm(mc["x_star"])),
       "objective": S.residuals(x, False)[1], "stats": st,
       "per_breakpoint_us": 1e6 * wall / max(st["breakpoints"], 1),
       "projection_factor_share": st["chol_ms"] * 1e-3 / wall, "gram_formation_share": st["gram_ms"] * 1e-3 / wall}
json.dump(res, open(out + ".json", "w"))
print(json.dumps(res))
S.close()
