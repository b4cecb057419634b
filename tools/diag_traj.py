import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import benlsip_b200 as B
from oracle import benlsip_oracle as O
from oracle.models import *

def cmp(tr_o, tr_g, nmax=40):
    print("outer", tr_o["outer_iters"], tr_g["outer_iters"], "inner", tr_o["inner_iters"], tr_g["stats"]["inner_iters"])
    for i, (a, b) in enumerate(zip(tr_o["inner"], tr_g["inner"])):
        if i >= nmax: break
        print(i, a["k"], b["k"], "mx %.15e %.15e" % (a["mx"], b["mx"]), "d %.6e %.6e" % (a["delta"], b["delta"]), "rho %.6e %.6e" % (a["rho"], b["rho"]),
              "pix %.6e %.6e" % (a["pix"], b["pix"]), "nfix", a["nb_fix"], b["nb_fix"], "ns %.6e %.6e" % (a["norm_s"], b["norm_s"]))

which = sys.argv[1]
if which == "sphere":
    P = SphereRegression
    tr_o, tr_g = {}, {}
    x_o, y_o = O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, max_outer_iter=100, max_inner_iter=250, trace=tr_o)
    x_g, y_g = B.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, max_outer_iter=100, max_inner_iter=250, trace=tr_g)
    print(x_o, y_o); print(x_g, y_g)
    cmp(tr_o, tr_g, 70)
    for a, b in zip(tr_o["outer"], tr_g["outer"]): print(a, b)
else:
    P = ExpSumProblem(4096, 16, seed=1)
    S = B.Solver(0); S.set_problem(P.M, P.n); S.use_builtin_model(B.MODEL_EXPSUM, 1e-3, 0.0, 1)
    tr_o, tr_g = {}, {}
    x_o, _ = O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, trace=tr_o)
    x_g, _ = B.tralcnllss(P.x0, None, None, None, None, None, None, None, None, solver=S, trace=tr_g)
    print(np.abs(x_o - x_g).max())
    cmp(tr_o, tr_g, 60)
