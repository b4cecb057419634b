"""Host-side set-up of the synthetic mixed-constraint family (BASELINE config[3]): the linear-equality matrix A, a feasible
start and the truth vector are O(m_lin * n) host data the CALLER supplies (as a BEnlsip.jl user would); residuals and the
Jacobian are generated on the device.  Same counter hash as the device generators (csrc/common.cuh)."""
from __future__ import annotations

import numpy as np

U32 = np.uint32


def _mix32(x):
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x ^= x >> U32(16)
        x *= U32(0x7FEB352D)
        x ^= x >> U32(15)
        x *= U32(0x846CA68B)
        x ^= x >> U32(16)
    return x


def _sym(seed, i, j):
    rk = _mix32(np.asarray(i, dtype=np.uint64).astype(np.uint32) ^ _mix32(U32(seed)))
    with np.errstate(over="ignore"):
        jj = np.asarray(j, dtype=np.uint64).astype(np.uint32) * U32(0x9E3779B9)
        h = _mix32(rk[:, None] + jj[None, :])
    return h.astype(np.float64) * 2.0 ** -31 - 1.0


def mixed_constraint_setup(n, m_lin, seed=5):
    """Returns dict(A, b, x0, x_star, rho2, xlow, xupp): A x0 = b, x_star on the same affine set, both inside [-1,1]^n."""
    j = np.arange(n)
    A = _sym(seed + 7, np.arange(m_lin), j)
    z0 = np.zeros(1)
    x_star = 0.6 * _sym(seed + 2, z0, j)[0]
    x_feas = 0.3 * _sym(seed + 3, z0, j)[0]
    x_star = x_star - A.T @ np.linalg.solve(A @ A.T, A @ (x_star - x_feas))
    return dict(A=A, b=A @ x_feas, x0=x_feas, x_star=x_star, rho2=float(x_star @ x_star), xlow=-np.ones(n), xupp=np.ones(n))
