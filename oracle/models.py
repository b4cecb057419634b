"""
ORACLE -- TEST INFRASTRUCTURE ONLY (see benlsip_oracle.py header).

NumPy statement of the synthetic problem families used by BASELINE.json's configs.  The
reference ships no large problems (only test/problems/sphere_regression.jl, restated here
verbatim as `SphereRegression`); cfg2..cfg5 are this build's definitions and
are generated identically on the device (benlsip.jl_b200/csrc/models.cu) from the same
counter-based 32-bit hash, so host oracle and CUDA library see the same data without ever
shipping it.

Hash ("lowbias32" finaliser, public domain):
    mix32(x): x ^= x>>16; x *= 0x7feb352d; x ^= x>>15; x *= 0x846ca68b; x ^= x>>16
    rowkey(seed,i) = mix32(uint32(i) ^ mix32(seed))
    h(seed,i,j)    = mix32(rowkey(seed,i) + uint32(j)*0x9E3779B9)
    unif(seed,i,j) = h * 2^-32  in [0,1)         sym(seed,i,j) = h * 2^-31 - 1  in [-1,1)
Both are exact in FP64.
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
_M1, _M2, _GOLD = U32(0x7FEB352D), U32(0x846CA68B), U32(0x9E3779B9)


def mix32(x):
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x ^= x >> U32(16)
        x *= _M1
        x ^= x >> U32(15)
        x *= _M2
        x ^= x >> U32(16)
    return x


def rowkey(seed, i):
    return mix32(np.asarray(i, dtype=np.uint64).astype(np.uint32) ^ mix32(U32(seed)))


def hash32(seed, i, j):
    """h(seed,i,j) broadcast over i (rows) x j (cols)."""
    rk = rowkey(seed, i)
    with np.errstate(over="ignore"):
        jj = np.asarray(j, dtype=np.uint64).astype(np.uint32) * _GOLD
        if rk.ndim == 1 and jj.ndim == 1:
            return mix32(rk[:, None] + jj[None, :])
        return mix32(rk + jj)


def unif(seed, i, j):
    return hash32(seed, i, j).astype(np.float64) * 2.0 ** -32


def sym(seed, i, j):
    return hash32(seed, i, j).astype(np.float64) * 2.0 ** -31 - 1.0


# --------------------------------------------------------------------------------------
class BoundProblem:
    """Common shape of a bound-only problem (m_A = 0, p = 0): A = zeros(0,n), c(x) = Float64[]."""

    M: int
    n: int

    def nlconstraints(self, x):
        return np.zeros(0)

    def jac_nlcons(self, x):
        return np.zeros((0, self.n))

    @property
    def A(self):
        return np.zeros((0, self.n))

    @property
    def b(self):
        return np.zeros(0)


# --------------------------------------------------------------------------------------
class GlmProblem(BoundProblem):
    """
    cfg3 / cfg5 family ("dense GLM-type residual", model id 1 in include/benlsip_b200.h):
        a_ij   = sym(seed_a, i, j) * cs_j,   cs_j = 10^(-cond_exp * j / n) / sqrt(n)
        phi(z) = z + 0.1 sin z
        r_i(x) = phi(a_i . x) - y_i,         J_ij = phi'(a_i . x) a_ij
        y_i    = phi(a_i . x_true) + noise * sym(seed_y, i, 0)
        x_true_j = 1.25*sgn_j if j % 10 == 0 else 0.9*sym(seed_x, 0, j);  box [-1,1]^n;  x0 = 0
    Rows [row0, row0 + M) of a global problem with M_total rows (row sharding).
    """

    def __init__(self, M, n, seed=3, noise=1e-3, cond_exp=0.0, row0=0, block=8192):
        self.M, self.n, self.seed, self.noise, self.cond_exp, self.row0 = int(M), int(n), int(seed), noise, cond_exp, int(row0)
        self.block = block
        j = np.arange(n)
        self.cs = glm_col_scale(n, cond_exp)
        self.x_true = glm_x_true(n, seed)
        self.xlow = -np.ones(n)
        self.xupp = np.ones(n)
        self.x0 = np.zeros(n)
        self._A = None
        self.y = self._gen_y()

    def rows(self, lo, hi):
        i = np.arange(self.row0 + lo, self.row0 + hi, dtype=np.uint64)
        return sym(self.seed, i, np.arange(self.n)) * self.cs[None, :]

    def design(self):
        if self._A is None:
            self._A = np.concatenate([self.rows(lo, min(lo + self.block, self.M)) for lo in range(0, self.M, self.block)])
        return self._A

    def _gen_y(self):
        z = self.design() @ self.x_true
        i = np.arange(self.row0, self.row0 + self.M, dtype=np.uint64)
        return z + 0.1 * np.sin(z) + self.noise * sym(self.seed + 1, i, np.zeros(1, dtype=np.uint64))[:, 0]

    def residuals(self, x):
        z = self.design() @ x
        return z + 0.1 * np.sin(z) - self.y

    def jac_res(self, x):
        z = self.design() @ x
        return (1.0 + 0.1 * np.cos(z))[:, None] * self.design()


def glm_col_scale(n, cond_exp):
    j = np.arange(n, dtype=np.float64)
    return np.power(10.0, -cond_exp * j / n) / np.sqrt(float(n))


def glm_x_true(n, seed):
    j = np.arange(n, dtype=np.uint64)
    s = sym(seed + 2, np.zeros(1, dtype=np.uint64), j)[0]
    xt = 0.9 * s
    on = (np.arange(n) % 10) == 0
    xt[on] = np.where(s[on] < 0, -1.25, 1.25)
    return xt


# --------------------------------------------------------------------------------------
class ExpSumProblem(BoundProblem):
    """
    cfg2 family ("exponential-sum data fit", model id 2): a sum of C = n/2 exponential decay channels,
    x = [a_0..a_{C-1}, b_0..b_{C-1}]; residual row i samples channel c = i mod C at time
    t_i = (floor(i/C) + 0.5) / ceil(M_total/C):
        model_i(x) = a_c exp(-b_c t_i),       r_i = model_i(x) - y_i
        J_{i,c} = exp(-b_c t_i),  J_{i,C+c} = -a_c t_i exp(-b_c t_i),  other columns 0
        (J is stored DENSE, M x n, as the reference's `Matrix{Float64}` would be)
        x_true: a_c = 1 + unif(seed,0,c), b_c = 0.5 + 3 unif(seed,1,c)
        y_i = model_i(x_true) + noise * sym(seed+1, i, 0)
        box = x_true +- 0.25, except every 8th parameter: true value exactly on its lower bound (box [xt, xt+0.5])
        x0 = box midpoint
    (Measured with this oracle: on a single K-term exponential sum observed on one time grid -- K = 128, and
    even K = 2 per channel -- the reference algorithm exhausts max_inner_iter in every outer iteration and never
    reaches its tolerance; the channel-separated sum is the variant it converges on: 8 outer / ~190 inner.)
    """

    def __init__(self, M, n, seed=1, noise=1e-3, row0=0, M_total=None):
        assert n % 2 == 0
        self.M, self.n, self.seed, self.noise, self.row0 = int(M), int(n), int(seed), noise, int(row0)
        self.M_total = int(M_total if M_total is not None else M)
        C = n // 2
        self.C = C
        self.x_true = expsum_x_true(n, seed)
        self.xlow = self.x_true - 0.25
        self.xupp = self.x_true + 0.25
        on = (np.arange(n) % 8) == 0
        self.xlow[on] = self.x_true[on]
        self.xupp[on] = self.x_true[on] + 0.5
        self.x0 = 0.5 * (self.xlow + self.xupp)
        i = np.arange(self.row0, self.row0 + self.M, dtype=np.int64)
        self.chan = i % C
        self.t = ((i // C).astype(np.float64) + 0.5) / float(-(-self.M_total // C))
        iu = i.astype(np.uint64)
        self.y = self._model(self.x_true) + noise * sym(seed + 1, iu, np.zeros(1, dtype=np.uint64))[:, 0]

    def _model(self, x):
        return x[self.chan] * np.exp(-x[self.C + self.chan] * self.t)

    def residuals(self, x):
        return self._model(x) - self.y

    def jac_res(self, x):
        C, c = self.C, self.chan
        e = np.exp(-x[C + c] * self.t)
        J = np.zeros((self.M, self.n))
        r = np.arange(self.M)
        J[r, c] = e
        J[r, C + c] = -(x[c] * self.t) * e
        return J


class DenseExpSumProblem(BoundProblem):
    """
    cfg2 as SURVEY 8d words it (model id 3): ONE sum of K = n/2 exponentials observed on one time grid,
        model_i(x) = sum_k a_k exp(-b_k t_i),   t_i = (i + 0.5) / M_total,   x = [a_0..a_{K-1}, b_0..b_{K-1}]
        J_{i,k} = exp(-b_k t_i),  J_{i,K+k} = -a_k t_i exp(-b_k t_i)
        x_true: a_k = 1 + unif(seed,0,k), b_k = 0.5 k + unif(seed,1,k);  y = model(x_true) + noise * sym(seed+1, i, 0)
        box = x_true +- 0.25, except every 8th parameter: true value exactly on its lower bound;  x0 = box midpoint
    Recovering 128 decay rates from one noisy curve is ill-posed: measured with this oracle, the reference algorithm exhausts
    max_inner_iter in every outer iteration on it (which is why BASELINE's cfg2 line uses the channel-separated family above).
    """

    def __init__(self, M, n, seed=1, noise=1e-3, row0=0, M_total=None):
        assert n % 2 == 0
        self.M, self.n, self.seed, self.noise, self.row0 = int(M), int(n), int(seed), noise, int(row0)
        self.M_total = int(M_total if M_total is not None else M)
        K = n // 2
        self.K = K
        k = np.arange(K, dtype=np.uint64)
        a = 1.0 + unif(seed, np.zeros(1, dtype=np.uint64), k)[0]
        b = 0.5 * np.arange(K, dtype=np.float64) + unif(seed, np.ones(1, dtype=np.uint64), k)[0]
        self.x_true = np.concatenate([a, b])
        self.xlow = self.x_true - 0.25
        self.xupp = self.x_true + 0.25
        on = (np.arange(n) % 8) == 0
        self.xlow[on] = self.x_true[on]
        self.xupp[on] = self.x_true[on] + 0.5
        self.x0 = 0.5 * (self.xlow + self.xupp)
        i = np.arange(self.row0, self.row0 + self.M, dtype=np.int64)
        self.t = (i.astype(np.float64) + 0.5) / float(self.M_total)
        self.y = self._model(self.x_true) + noise * sym(seed + 1, i.astype(np.uint64), np.zeros(1, dtype=np.uint64))[:, 0]

    def _model(self, x):
        return np.exp(-np.outer(self.t, x[self.K:])) @ x[:self.K]

    def residuals(self, x):
        return self._model(x) - self.y

    def jac_res(self, x):
        E = np.exp(-np.outer(self.t, x[self.K:]))
        return np.hstack([E, -(self.t[:, None] * x[None, :self.K]) * E])


def expsum_x_true(n, seed):
    C = n // 2
    c = np.arange(C, dtype=np.uint64)
    a = 1.0 + unif(seed, np.zeros(1, dtype=np.uint64), c)[0]
    b = 0.5 + 3.0 * unif(seed, np.ones(1, dtype=np.uint64), c)[0]
    return np.concatenate([a, b])


# --------------------------------------------------------------------------------------
class SphereRegression:
    """cfg1 -- test/problems/sphere_regression.jl:10-32, verbatim (M=4, n=3, m_A=1, p=1)."""

    M, n = 4, 3
    xlow = np.array([-2.0, -1.5, 0.0])
    xupp = np.array([2.0, 1.5, 2.0])
    A = np.array([[1.0, 2.0, -1.0]])
    b = np.array([0.5])
    x0 = np.array([1.0, 0.5, 1.5])

    @staticmethod
    def residuals(x):
        return np.array([
            x[0] ** 2 + x[1] ** 2 - 2 * x[0] + np.sin(x[0] + x[1]) - 1.5,
            x[0] * x[1] + 0.5 * np.cos(2 * x[0]) - 0.8,
            (x[0] - 1.0) ** 2 + (x[1] - 0.5) ** 2 - x[2],
            x[2] ** 2 - x[0] + 0.3 * np.sin(x[2]) - 0.2,
        ])

    @staticmethod
    def jac_res(x):
        return np.array([
            [2 * x[0] - 2 + np.cos(x[0] + x[1]), 2 * x[1] + np.cos(x[0] + x[1]), 0.0],
            [x[1] - np.sin(2 * x[0]), x[0], 0.0],
            [2 * (x[0] - 1), 2 * (x[1] - 0.5), -1.0],
            [-1.0, 0.0, 2 * x[2] + 0.3 * np.cos(x[2])],
        ])

    @staticmethod
    def nlconstraints(x):
        return np.array([x[0] ** 2 + x[1] ** 2 + x[2] ** 2 - 3])

    @staticmethod
    def jac_nlcons(x):
        return np.array([[2 * x[0], 2 * x[1], 2 * x[2]]])


# --------------------------------------------------------------------------------------
class MixedConstraintProblem:
    """
    cfg4 family, host-callback form (shrunk): GLM residual r_i = phi(a_i.x) - y_i with
      * m_lin linear equalities A x = b (A from the hash, b = A x_feas, x0 = x_feas strictly inside the box),
      * p = 1 nonlinear equality  ||x||^2 = ||x_star||^2  (sphere through a box-feasible point),
      * box [-1, 1]^n.
    Exercises the general projection (Cholesky of A~A~', triangular solves), the AL multiplier / penalty updates and
    the C-block of the GN Hessian.  Dense NumPy callbacks (the reference's calling convention).
    """

    def __init__(self, M, n, m_lin, seed=5, noise=1e-3):
        self.M, self.n, self.m_lin, self.seed = int(M), int(n), int(m_lin), int(seed)
        i = np.arange(M, dtype=np.uint64)
        j = np.arange(n)
        self.Ad = sym(seed, i, j) * glm_col_scale(n, 0.0)[None, :]  # == the device GLM design with the same seed
        self.A = sym(seed + 7, np.arange(m_lin, dtype=np.uint64), j)
        z0 = np.zeros(1, dtype=np.uint64)
        self.x_star = 0.6 * sym(seed + 2, z0, j)[0]
        x_feas = 0.3 * sym(seed + 3, z0, j)[0]
        # put x_star on the affine set through x_feas:  x_star <- x_star - A^+ A (x_star - x_feas)
        AAt = self.A @ self.A.T
        self.x_star = self.x_star - self.A.T @ np.linalg.solve(AAt, self.A @ (self.x_star - x_feas))
        self.b = self.A @ x_feas
        self.x0 = x_feas
        self.rho2 = float(self.x_star @ self.x_star)
        self.xlow = -np.ones(n)
        self.xupp = np.ones(n)
        zt = self.Ad @ self.x_star
        self.y = zt + 0.1 * np.sin(zt) + noise * sym(seed + 1, i, z0)[:, 0]

    def residuals(self, x):
        z = self.Ad @ x
        return z + 0.1 * np.sin(z) - self.y

    def jac_res(self, x):
        z = self.Ad @ x
        return (1.0 + 0.1 * np.cos(z))[:, None] * self.Ad

    def nlconstraints(self, x):
        return np.array([x @ x - self.rho2])

    def jac_nlcons(self, x):
        return (2.0 * x)[None, :]
