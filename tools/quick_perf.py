"""Quick kernel-class timing on one GPU (CUDA events inside the library): python tools/quick_perf.py [M n]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import benlsip_b200 as B

cases = [(1_000_000, 256), (2_000_000, 1024), (500_000, 2048), (250_000, 4096)]
if len(sys.argv) >= 3:
    cases = [(int(sys.argv[1]), int(sys.argv[2]))]
S = B.Solver(0)
print(S.device_info())
for M, n in cases:
    S.set_problem(M, n)
    S.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, 3)
    x0 = S.model_vectors()["x0"]
    S.eval_jacobian(x0 + 0.1)
    S.residuals(x0 + 0.1, False)
    out = {"M": M, "n": n}
    for kind, name in [(0, "jtjv"), (1, "jv"), (2, "jtw"), (3, "residual"), (4, "jacobian")] + ([(6, "jacobian_fused_jtr")] if n <= 1024 else []):
        ms, nbytes = S.time_kernel(kind, 20)
        out[name] = {"ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1)}
    ms, fl = S.time_kernel(5, 3)
    out["gram"] = {"ms": round(ms, 3), "TFLOPs_issued": round(fl / ms / 1e9, 2), "TFLOPs_2Mn2": round(2.0 * M * n * n / ms / 1e9, 2)}
    print(json.dumps(out), flush=True)
S.close()
