"""Bound-constrained dense NLS with a device-side model (BASELINE config[2] family, shrunk): residuals and the Jacobian are
generated on the GPU, the outer loop stays on the host.  python examples/glm_device_model.py [M n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benlsip_b200 as B

M, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1_000_000, 1024)
S = B.Solver(0)
S.set_problem(M, n)
S.use_builtin_model(B.MODEL_GLM, noise=1e-3, cond_exp=0.0, seed=3)
x0 = S.model_vectors()["x0"]
for name, setup in [("literal (reference semantics)", lambda: None),
                    ("gram-apply", lambda: S.set_hessian_mode(B.HESSIAN_GRAM)),
                    ("incremental cauchy", lambda: (S.set_hessian_mode(B.HESSIAN_MATRIX_FREE), S.set_cauchy_mode(B.CAUCHY_INCREMENTAL)))]:
    setup()
    S.reset_stats()
    tr = {}
    t0 = time.perf_counter()
    x, _ = B.tralcnllss(x0, None, None, None, None, None, None, None, None, solver=S, trace=tr)
    st = tr["stats"]
    print(f"{name:32s} {time.perf_counter() - t0:7.3f} s  outer {tr['outer_iters']} inner {st['inner_iters']} "
          f"applies {st['hess_mul']} J passes {st['j_passes']} objective {S.residuals(x, False)[1]:.6e}")
S.close()
