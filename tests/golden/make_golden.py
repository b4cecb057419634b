"""Generates tests/golden/*.json from the oracle (restatement-derived goldens: the reference ships no trajectory and Julia
is not available here -- see oracle/benlsip_oracle.py header).  python tests/golden/make_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np

from oracle import benlsip_oracle as O
from oracle.models import GlmProblem, MixedConstraintProblem, SphereRegression


def run(P, **kw):
    tr = {}
    x, y = O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, trace=tr, **kw)
    return dict(x=x.tolist(), y=np.asarray(y).tolist(), outer_iters=tr["outer_iters"], inner_iters=tr["inner_iters"],
                minor_iters=tr.get("minor_iters", 0), cg_iters=tr.get("cg_iters", 0), breakpoints=tr.get("breakpoints", 0),
                mu=tr["mu"], fixvars_words=[int(w) for w in tr["fixvars_words"]],
                objective=float(np.sum(P.residuals(x) ** 2)),
                inner=[{k: r[k] for k in ("k", "mx", "delta", "pix", "nb_fix", "rho", "pred", "norm_s", "omega_tol", "bp_cum", "cg_cum")}
                       for r in tr["inner"]])


cases = {
    "glm_4096_64": (GlmProblem(4096, 64, seed=3), {}),
    "glm_20000_256": (GlmProblem(20000, 256, seed=3), {}),
    "glm_6000_1024": (GlmProblem(6000, 1024, seed=3), {}),
    "mixed_600_24_4": (MixedConstraintProblem(600, 24, 4), dict(max_outer_iter=60, max_inner_iter=200)),
    "sphere_regression": (SphereRegression, dict(max_outer_iter=100, max_inner_iter=250)),
}
for name, (P, kw) in cases.items():
    json.dump(run(P, **kw), open(os.path.join(HERE, name + ".json"), "w"))
    print("wrote", name)
# cfg5 family (ill-conditioned): twelve consecutive inner steps from x0 (Cauchy point + projected CG), per-step quantities
P = GlmProblem(3000, 96, seed=3, cond_exp=6.0)
L0 = O._cholesky_lower(np.zeros((0, 0)))
cons = O.MixedConstraints(P.A, L0, l=P.xlow, u=P.xupp)
x = P.x0.copy()
steps = []
for it in range(12):
    J, r = P.jac_res(x), P.residuals(x)
    g = J.T @ r
    H = O.AlHessian(J, np.zeros((0, 96)), 0.0)
    delta = 0.1 * np.linalg.norm(g)
    tr = {}
    s, pred = O.inner_step(x, g, H, L0, cons, delta, 50, 0.1, 0.1, trace=tr)
    steps.append(dict(x=x.tolist(), g=g.tolist(), delta=delta, s=s.tolist(), pred=pred, cg_iters=tr.get("cg_iters", 0),
                      breakpoints=tr.get("breakpoints", 0), minor_iters=tr.get("minor_iters", 0),
                      fixvars_words=[int(w) for w in cons.fixvars_words()]))
    x = x + s
json.dump(dict(M=3000, n=96, cond_exp=6.0, steps=steps), open(os.path.join(HERE, "glm_cfg5_3000_96_steps.json"), "w"))
print("wrote glm_cfg5_3000_96_steps")
# HS48 projection (the reference's literal golden vector, test/structures.jl:37-58)
json.dump(dict(A=[[1.0, 1, 1, 1, 1], [0, 0, 1, -2, -2]], x=[3.0, 5, -3, 2, -2], fixed=[0, 1], projection=[0.0, 0, 0, 2, -2]),
          open(os.path.join(HERE, "hs48_projection.json"), "w"))
