"""Generates the headline-regime golden (GLM family, n = 1024, M/n >> 1: Cauchy breakpoints dominate, projected CG does not
run -- the regime of BASELINE cfg3) from the oracle, once, on the host.  Takes ~10-20 minutes of CPU (thousands of passes
over a 1.6 GB Jacobian), which is why it is a committed fixture and not a live comparison.
    python tests/golden/make_golden_headline.py [M] [n]"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np

from oracle import benlsip_oracle as O
from oracle.models import GlmProblem

M = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
P = GlmProblem(M, n, seed=3)
tr = {}
t0 = time.time()
x, y = O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, trace=tr)
out = dict(M=M, n=n, seed=3, x=x.tolist(), outer_iters=tr["outer_iters"], inner_iters=tr["inner_iters"],
           minor_iters=tr.get("minor_iters", 0), cg_iters=tr.get("cg_iters", 0), breakpoints=tr.get("breakpoints", 0),
           mu=tr["mu"], fixvars_words=[int(w) for w in tr["fixvars_words"]], counters=tr["counters"],
           objective=float(np.sum(P.residuals(x) ** 2)),
           inner=[{k: r[k] for k in ("k", "mx", "delta", "pix", "nb_fix", "rho", "pred", "norm_s", "omega_tol", "bp_cum", "cg_cum")}
                  for r in tr["inner"]],
           outer=tr["outer"], seconds=time.time() - t0)
json.dump(out, open(os.path.join(HERE, f"glm_{M}_{n}.json"), "w"))
print("wrote", f"glm_{M}_{n}.json", "outer", out["outer_iters"], "inner", out["inner_iters"], "bp", out["breakpoints"], "cg",
      out["cg_iters"], "seconds", out["seconds"])
