import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
import benlsip_b200 as B
S = B.Solver(0)
M, n = 10_000_000, 1024
S.set_problem(M, n); S.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, 3)
x0 = S.model_vectors()["x0"]
tr = {}
x, _ = B.tralcnllss(x0, None, None, None, None, None, None, None, None, solver=S, trace=tr)
prev = 0
for r in tr["inner"]:
    print(r["k"], "bp", r["bp_cum"] - prev, "nb_fix", r["nb_fix"], "delta %.3e" % r["delta"], "rho %.3g" % r["rho"], "norm_s %.2e" % r["norm_s"], "pix %.2e" % r["pix"])
    prev = r["bp_cum"]
