"""Small end-to-end exercise of every kernel family for compute-sanitizer memcheck:
compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import benlsip_b200 as B
from oracle.models import MixedConstraintProblem, SphereRegression

S = B.Solver(0)
rng = np.random.default_rng(0)
for M, n in [(37, 5), (300, 130), (260, 1000), (70, 2048), (40, 4096)]:
    J = rng.standard_normal((M, n))
    S.set_problem(M, n)
    S.upload_jacobian(J)
    v = rng.standard_normal(n)
    assert np.allclose(S.hess_mul(v), J.T @ (J @ v))
    S.vthv(v), S.jv(v), S.jtw(rng.standard_normal(M))
    G, _ = S.gram()
    assert np.allclose(G, J.T @ J)
for mid, n in [(B.MODEL_GLM, 96), (B.MODEL_EXPSUM, 16)]:
    S.set_problem(2000, n)
    S.use_builtin_model(mid, 1e-3, 0.0, 3)
    x0 = S.model_vectors()["x0"]
    B.tralcnllss(x0, None, None, None, None, None, None, None, None, solver=S, max_outer_iter=3, max_inner_iter=5)
    S.set_hessian_mode(B.HESSIAN_GRAM)
    B.tralcnllss(x0, None, None, None, None, None, None, None, None, solver=S, max_outer_iter=2, max_inner_iter=3)
    S.set_hessian_mode(B.HESSIAN_MATRIX_FREE)
S.close()
P = SphereRegression
B.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, max_outer_iter=3, max_inner_iter=10)
P = MixedConstraintProblem(200, 12, 3)
B.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, max_outer_iter=3, max_inner_iter=10)
print("sanitize_small ok")
