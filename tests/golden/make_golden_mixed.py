"""Generates tests/golden/mixed_<M>_<n>_<m>.json: the cfg4 family (linear equalities + sphere constraint + box) at a mid size, from
the oracle (literal block Cholesky factor rebuilt per breakpoint, as the reference does).  (4000, 64, 8) takes ~3.5 minutes of CPU.
    python tests/golden/make_golden_mixed.py [M n m_lin]"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np

from oracle import benlsip_oracle as O
from oracle.models import MixedConstraintProblem

M, n, m = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4000, 64, 8)
P = MixedConstraintProblem(M, n, m)
tr = {}
t0 = time.time()
x, y = O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, trace=tr,
                    max_outer_iter=60, max_inner_iter=200)
print("seconds", time.time() - t0, "outer", tr["outer_iters"], "inner", tr["inner_iters"], "minor", tr.get("minor_iters"), "cg",
      tr.get("cg_iters"), "bp", tr.get("breakpoints"), "mu", tr["mu"], flush=True)
out = dict(M=M, n=n, m_lin=m, x=x.tolist(), y=np.asarray(y).tolist(), outer_iters=tr["outer_iters"], inner_iters=tr["inner_iters"],
           minor_iters=tr.get("minor_iters", 0), cg_iters=tr.get("cg_iters", 0), breakpoints=tr.get("breakpoints", 0), mu=tr["mu"],
           fixvars_words=[int(w) for w in tr["fixvars_words"]], objective=float(np.sum(P.residuals(x) ** 2)),
           inner=[{k: r[k] for k in ("k", "mx", "delta", "pix", "nb_fix", "rho", "pred", "norm_s", "omega_tol", "bp_cum", "cg_cum")}
                  for r in tr["inner"]])
json.dump(out, open(os.path.join(HERE, f"mixed_{M}_{n}_{m}.json"), "w"))
