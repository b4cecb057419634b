"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Literal FP64 NumPy restatement of the live hot path of pierre-borie/BEnlsip.jl (the inner
Gauss-Newton trust-region subproblem solve and the outer augmented-Lagrangian driver that
calls it).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the CUDA library never does.

PARITY STATUS: the reference's own tests pin (i) one literal golden vector (HS48 projection,
test/structures.jl:37-58), (ii) structural identities (test/structures.jl:1-35, :60-78) and
(iii) end-state inequalities on sphere_regression (test/problems/sphere_regression.jl:63-65).
All of those are checked in tests/test_oracle_reference_fixtures.py.  Iteration counts and
intermediate iterates are NOT pinned by the reference (no Julia in this image, no trace
shipped) => "trajectory parity unpinned": trajectories are pinned by this restatement only.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Operation order is kept literal (e.g. `H*s+g`, `dot(s_c,Hd)+dot(g,d)`, mu folded into the C
gemv) and the reference's quirks (SURVEY.md section 8a traps T1-T9) are reproduced, not fixed.
Indices are 0-based here; `fixvars` is a bool array (Julia BitVector).
"""
from __future__ import annotations

import math
import numpy as np

try:  # scipy is only used for triangular solves; tiny fallback below keeps the oracle standalone
    from scipy.linalg import solve_triangular as _solve_tri
except Exception:  # pragma: no cover
    _solve_tri = None

SQRT_EPS = math.sqrt(np.finfo(np.float64).eps)  # sqrt(eps(Float64)) = 1.4901161193847656e-08

# CG_status enum, src/basic_tralcnlss.jl:12
SOLVED, BOUND_HIT, NEGATIVE_CURVATURE, MAX_ITER_REACHED = 0, 1, 2, 3


class PosDefException(ArithmeticError):
    """Julia LinearAlgebra.PosDefException (raised by `cholesky`)."""


def _cholesky_lower(Mx: np.ndarray) -> np.ndarray:
    """Julia `cholesky(M).L`.  0x0 is legal (bound-only problems, SURVEY 8a 'degenerate shapes')."""
    Mx = np.asarray(Mx, dtype=np.float64)
    if Mx.shape[0] == 0:
        return np.zeros((0, 0))
    try:
        return np.linalg.cholesky(Mx)
    except np.linalg.LinAlgError as e:  # same failure mode as the reference
        raise PosDefException(str(e))


def _trisolve(L: np.ndarray, b: np.ndarray, lower: bool) -> np.ndarray:
    if L.shape[0] == 0:
        return np.zeros_like(b, dtype=np.float64)
    if _solve_tri is not None:
        return _solve_tri(L, b, lower=lower, check_finite=False)
    return np.linalg.solve(L, b)  # pragma: no cover


# --------------------------------------------------------------------------------------
# AlHessian  (src/basic_tralcnlss.jl:6-10, :92-96, :102-106)
# --------------------------------------------------------------------------------------
class AlHessian:
    """Matrix-free GN Hessian H = J'J + mu C'C.  src/basic_tralcnlss.jl:6-10."""

    def __init__(self, J, C, mu, counters=None):
        self.J = J
        self.C = C
        self.mu = float(mu)
        self.counters = counters  # optional dict for J-pass accounting

    def _count(self, key, k=1):
        if self.counters is not None:
            self.counters[key] = self.counters.get(key, 0) + k

    def mul(self, v):
        """`H*v`, src/basic_tralcnlss.jl:102-106: Jv = J*v; muCv = mu*C*v; J'Jv + C'muCv."""
        self._count("hess_mul")
        self._count("jv")
        self._count("jtw")
        Jv = self.J @ v
        muCv = (self.mu * self.C) @ v  # Julia parses H.mu*H.C*v as (mu*C)*v
        return self.J.T @ Jv + self.C.T @ muCv

    def vthv(self, v):
        """`vthv(H,v)`, src/basic_tralcnlss.jl:92-96."""
        self._count("vthv")
        self._count("jv")
        Jv = self.J @ v
        Cv = self.C @ v
        return float(np.dot(Jv, Jv) + self.mu * np.dot(Cv, Cv))


# --------------------------------------------------------------------------------------
# MixedConstraints  (src/polyhedral_constraints.jl)
# --------------------------------------------------------------------------------------
class MixedConstraints:
    """src/polyhedral_constraints.jl:1-7.  `chol` is stored as its lower factor L."""

    def __init__(self, A, chol_aat_L, l=None, u=None, fixed=None):
        A = np.asarray(A, dtype=np.float64)
        n = A.shape[1]
        self.lineq = A
        self.xlow = np.full(n, -np.inf) if l is None else np.asarray(l, dtype=np.float64)
        self.xupp = np.full(n, np.inf) if u is None else np.asarray(u, dtype=np.float64)
        if fixed is None:  # ctor :9-18
            self.fixvars = np.zeros(n, dtype=bool)
            self.chol = chol_aat_L
        else:  # ctor :20-29
            self.fixvars = np.asarray(fixed, dtype=bool).copy()
            self.chol = cholesky_aug_aat(A, self.fixvars, chol_aat_L) if self.fixvars.any() else chol_aat_L
        self.n_chol_rebuilds = 0

    def nb_fix(self):
        """src/polyhedral_constraints.jl:31."""
        return int(np.count_nonzero(self.fixvars))

    def fixvars_words(self):
        """Julia BitVector.chunks layout: bit (i)&63 of word (i)>>6, LSB first (0-based i)."""
        n = self.fixvars.shape[0]
        nw = (n + 63) // 64
        bits = np.zeros(nw * 64, dtype=np.uint8)
        bits[:n] = self.fixvars
        return np.packbits(bits.reshape(nw, 64), axis=1, bitorder="little").view(np.uint64).reshape(nw)


def cholesky_aug_aat(A, fix_bounds, chol_aat_L):
    """src/polyhedral_constraints.jl:35-59.  L = [L_A 0; G' chol(I - G'G)], G = L_A \\ A[:,fix]."""
    m, n = A.shape
    p = int(np.count_nonzero(fix_bounds))
    mpp = m + p
    assert mpp <= n  # :43
    Hm = np.eye(p)
    L = np.zeros((mpp, mpp))
    A_act_cols = A[:, fix_bounds]
    G = _trisolve(chol_aat_L, A_act_cols, lower=True) if m > 0 else np.zeros((0, p))
    Hm = Hm - G.T @ G  # mul!(H, G', G, -1, 1), :52
    L[:m, :m] = chol_aat_L
    L[m:, :m] = G.T
    L[m:, m:] = _cholesky_lower(Hm)  # :57 (PosDefException possible)
    return L


def update_chol(lincons: MixedConstraints, chol_aat_L):
    """src/polyhedral_constraints.jl:62-68 (rebuild from scratch on every call)."""
    lincons.chol = cholesky_aug_aat(lincons.lineq, lincons.fixvars, chol_aat_L)
    lincons.n_chol_rebuilds += 1


def left_mul_tr(lincons: MixedConstraints, y):
    """src/polyhedral_constraints.jl:72-84:  A~' y."""
    m, n = lincons.lineq.shape
    if lincons.fixvars.any():
        x = lincons.lineq.T @ y[:m]
        x[lincons.fixvars] += y[m:]
    else:
        x = lincons.lineq.T @ y
    return x


def left_mul(lincons: MixedConstraints, x):
    """src/polyhedral_constraints.jl:86-98:  A~ x = [A x; x[fix]]."""
    m, _ = lincons.lineq.shape
    y = np.empty(m + lincons.nb_fix())
    if lincons.fixvars.any():
        y[:m] = lincons.lineq @ x
        y[m:] = x[lincons.fixvars]
    else:
        y[:] = lincons.lineq @ x
    return y


def projection_nullspace(lincons: MixedConstraints, r):
    """src/polyhedral_constraints.jl:104-118."""
    assert not lincons.fixvars.any()  # :110
    y = _trisolve(lincons.chol, lincons.lineq @ r, lower=True)
    w = _trisolve(lincons.chol.T, y, lower=False)
    return r - lincons.lineq.T @ w


def projection_subspace(lincons: MixedConstraints, r):
    """src/polyhedral_constraints.jl:120-136."""
    m, n = lincons.lineq.shape
    mpp = m + lincons.nb_fix()
    assert m < mpp <= n  # :128
    y = _trisolve(lincons.chol, left_mul(lincons, r), lower=True)
    w = _trisolve(lincons.chol.T, y, lower=False)
    return r - left_mul_tr(lincons, w)


def projection(lincons: MixedConstraints, r):
    """src/polyhedral_constraints.jl:150-170 (`projection` and `projection!`)."""
    if lincons.fixvars.any():
        return projection_subspace(lincons, r)
    return projection_nullspace(lincons, r)


def active_bounds_reset(lincons: MixedConstraints, x, chol_aat_L, atol=SQRT_EPS):
    """`active_bounds!`, src/polyhedral_constraints.jl:203-215 -- OVERWRITES fixvars from x."""
    lincons.fixvars[:] = ((x - lincons.xlow) <= atol) | ((lincons.xupp - x) <= atol)
    update_chol(lincons, chol_aat_L)


def active_bounds(lincons: MixedConstraints, x, s, delta, atol=SQRT_EPS):
    """`active_bounds`, src/polyhedral_constraints.jl:219-237.  Returns ascending indices (0-based)."""
    s_l = np.maximum(lincons.xlow - x, -delta)
    s_u = np.minimum(lincons.xupp - x, delta)
    at_bound = ((s - s_l) <= atol) | ((s_u - s) <= atol)
    return np.flatnonzero(at_bound)


def add_active(lincons: MixedConstraints, chol_aat_L, ind):
    """`add_active!` (Int and Vector{Int} methods), src/polyhedral_constraints.jl:240-261."""
    if np.isscalar(ind) or isinstance(ind, (int, np.integer)):
        if ind < 0:
            raise IndexError("BoundsError: add_active! with ind=-1 (next_breakpoint found no breakpoint)")
        lincons.fixvars[int(ind)] = True
    else:
        lincons.fixvars[np.asarray(ind, dtype=np.int64)] = True
    update_chol(lincons, chol_aat_L)


# --------------------------------------------------------------------------------------
# Function / derivative evaluation  (src/basic_tralcnlss.jl:32-85)
# --------------------------------------------------------------------------------------
def new_point(x, y, mu, residuals, nlconstraints, jac_res, jac_nlcons, counters=None):
    """src/basic_tralcnlss.jl:32-49."""
    rx, cx = residuals(x), nlconstraints(x)
    Jx, Cx = jac_res(x), jac_nlcons(x)
    y_bar = y + mu * cx
    mx = 0.5 * np.dot(rx, rx) + np.dot(y, cx) + 0.5 * mu * np.dot(cx, cx)
    g = Jx.T @ rx + Cx.T @ y_bar
    if counters is not None:
        counters["res_eval"] = counters.get("res_eval", 0) + 1
        counters["jac_eval"] = counters.get("jac_eval", 0) + 1
        counters["jtw"] = counters.get("jtw", 0) + 1
    H = AlHessian(Jx, Cx, mu, counters)
    return rx, cx, y_bar, float(mx), g, H


def evaluate_al(x, y, mu, residuals, nlconstraints, counters=None):
    """src/basic_tralcnlss.jl:51-61."""
    rx, cx = residuals(x), nlconstraints(x)
    mx = 0.5 * np.dot(rx, rx) + np.dot(y, cx) + 0.5 * mu * np.dot(cx, cx)
    if counters is not None:
        counters["res_eval"] = counters.get("res_eval", 0) + 1
    return rx, cx, float(mx)


def first_derivatives(x, y, mu, rx, cx, jac_res, jac_nlcons, counters=None):
    """src/basic_tralcnlss.jl:63-77."""
    Jx, Cx = jac_res(x), jac_nlcons(x)
    y_bar = y + mu * cx
    g = Jx.T @ rx + Cx.T @ y_bar
    if counters is not None:
        counters["jac_eval"] = counters.get("jac_eval", 0) + 1
        counters["jtw"] = counters.get("jtw", 0) + 1
    return y_bar, Jx, Cx, g


# --------------------------------------------------------------------------------------
# Small helpers  (src/basic_tralcnlss.jl:153-163, :793-844, :869-911)
# --------------------------------------------------------------------------------------
def initial_tolerances(mu, omega0, eta0, k_crit, k_feas):
    """src/basic_tralcnlss.jl:153-163."""
    return omega0 / (mu ** k_crit), eta0 / (mu ** k_feas)


def factor_to_boundary(p, w, w_l, w_u, atol=1e-10):
    """src/basic_tralcnlss.jl:793-809 (note: loops over ALL i, no fixvars test)."""
    gamma = np.inf
    with np.errstate(all="ignore"):
        neg = p <= -atol
        pos = p >= atol
        if neg.any():
            gamma = min(gamma, float(np.min((w_l[neg] - w[neg]) / p[neg])))
        if pos.any():
            gamma = min(gamma, float(np.min((w_u[pos] - w[pos]) / p[pos])))
    return gamma


def initial_tr(g, tr_factor=0.1):
    """src/basic_tralcnlss.jl:817-819."""
    return tr_factor * float(np.linalg.norm(g))


def update_tr(delta, rho, eta1, eta2, gamma1, gamma2):
    """src/basic_tralcnlss.jl:821-837 (NaN rho => unchanged, trap T8)."""
    if rho > eta2:
        return gamma2 * delta
    elif rho < eta1:
        return gamma1 * delta
    return delta


def norm_reduced_gradient(g, polyhedron: MixedConstraints):
    """src/basic_tralcnlss.jl:869-875."""
    return float(np.linalg.norm(projection(polyhedron, -g)))


def criticality_measure(g, lincons: MixedConstraints):
    """src/basic_tralcnlss.jl:839-844."""
    return norm_reduced_gradient(g, lincons)


def least_squares_multipliers(x, residuals, jac_res, jac_nlcons):
    """src/basic_tralcnlss.jl:887-903:  y = -(CC')^{-1} C J'r."""
    g = jac_res(x).T @ residuals(x)
    C = jac_nlcons(x)
    L = _cholesky_lower(C @ C.T)
    b = -(C @ g)
    v = _trisolve(L, b, lower=True)
    return _trisolve(L.T, v, lower=False)


def first_order_multipliers(y, cx, mu):
    """src/basic_tralcnlss.jl:905-911."""
    return y + mu * cx


# --------------------------------------------------------------------------------------
# Cauchy step  (src/basic_tralcnlss.jl:536-562, :574-639)
# --------------------------------------------------------------------------------------
def next_breakpoint(d, s, d_l, d_u, fix_bounds):
    """src/basic_tralcnlss.jl:536-562.  Strict `<` => lowest index wins ties; ind=-1 if none."""
    n = d.shape[0]
    with np.errstate(all="ignore"):
        theta_try = np.full(n, np.inf)
        neg = (d < 0) & ~fix_bounds
        pos = (d > 0) & ~fix_bounds
        theta_try[neg] = (d_l[neg] - s[neg]) / d[neg]
        theta_try[pos] = (d_u[pos] - s[pos]) / d[pos]
    theta, ind = np.inf, -1
    free = ~fix_bounds
    if free.any():
        # sequential strict-< scan == first occurrence of the minimum among candidates < Inf
        # (NaN candidates never satisfy `<`, mirror that)
        cand = np.where(free & ~np.isnan(theta_try), theta_try, np.inf)
        j = int(np.argmin(cand))
        if cand[j] < np.inf:
            theta, ind = float(cand[j]), j
    return theta, ind


def cauchy_step(x, g, H: AlHessian, chol_aat_L, lincons: MixedConstraints, delta, trace=None):
    """6-argument (live) `cauchy_step`, src/basic_tralcnlss.jl:574-639."""
    m, n = lincons.lineq.shape
    nmm = n - m
    s_c = np.zeros(n)

    active_bounds_reset(lincons, x, chol_aat_L)  # :591
    d = projection(lincons, -g)  # :592
    d_u = np.minimum(lincons.xupp - x, delta)  # :602
    d_l = np.maximum(lincons.xlow - x, -delta)  # :603

    Hd = H.mul(d)  # :609
    phi_p = np.dot(s_c, Hd) + np.dot(g, d)  # :610
    phi_pp = np.dot(d, Hd)  # :611
    min_found = False
    n_break = 0
    while (not min_found) and (lincons.nb_fix() < nmm):  # :615
        theta, ind = next_breakpoint(d, s_c, d_l, d_u, lincons.fixvars)  # :617
        delta_t = (-phi_p / phi_pp) if phi_pp > 0 else 0.0  # :618
        if phi_p >= 0:  # :620
            min_found = True
        elif phi_p < 0 and phi_pp > 0 and delta_t < theta:  # :622
            delta_t = -phi_p / phi_pp
            s_c = s_c + delta_t * d  # :625
            min_found = True
        else:  # :627-635
            s_c = s_c + theta * d
            add_active(lincons, chol_aat_L, ind)
            d = projection(lincons, -g)
            Hd = H.mul(d)
            phi_p = np.dot(s_c, Hd) + np.dot(g, d)
            phi_pp = np.dot(d, Hd)
            n_break += 1
    if trace is not None:
        trace["breakpoints"] = trace.get("breakpoints", 0) + n_break
    return s_c


# --------------------------------------------------------------------------------------
# Minor iterate: projected CG + linesearch  (src/basic_tralcnlss.jl:649-791)
# --------------------------------------------------------------------------------------
def projected_cg(g_minor, H: AlHessian, w_l, w_u, lincons: MixedConstraints, kappa2, atol=SQRT_EPS, trace=None):
    """src/basic_tralcnlss.jl:690-764.  Returns (w, status) with status None per trap T3."""
    m, n = lincons.lineq.shape
    w = np.zeros(n)
    r = np.array(g_minor, dtype=np.float64, copy=True)
    v = projection(lincons, r)
    rtv = np.dot(r, v)
    p = -v
    tol_cg = kappa2 * float(np.linalg.norm(v))
    tol_negcurve = atol
    it = 1
    max_iter = 2 * (n - m - lincons.nb_fix())
    approx_solved = False
    neg_curvature = False
    outside_region = False
    n_hp = 0
    while (not approx_solved) and (not outside_region) and (not neg_curvature) and it <= max_iter:
        Hp = H.mul(p)  # :722
        n_hp += 1
        pHp = np.dot(p, Hp)  # :723
        if pHp <= tol_negcurve:  # :725
            neg_curvature = True
            if abs(pHp) > tol_negcurve:  # dead for PSD H (trap T2)
                gamma = factor_to_boundary(p, w, w_l, w_u)
                w = w + gamma * p
        else:
            rtv = np.dot(r, v)  # :732
            alpha = rtv / pHp
            gamma = factor_to_boundary(p, w, w_l, w_u)
            outside_region = alpha > gamma
            if outside_region:
                w = w + gamma * p
            else:
                w = w + alpha * p
                r = r + alpha * Hp
                v = projection(lincons, r)  # :741
                rtv_next = np.dot(r, v)
                beta = rtv_next / rtv
                p = -v + beta * p  # axpby!(-1, v, beta, p), :745
                rtv = rtv_next
                approx_solved = abs(rtv) < tol_cg
                it += 1
    if approx_solved:
        status = SOLVED
    elif outside_region:
        status = BOUND_HIT
    elif neg_curvature:
        status = NEGATIVE_CURVATURE
    elif it == max_iter:
        status = MAX_ITER_REACHED
    else:
        status = None  # trap T3
    if trace is not None:
        trace["cg_iters"] = trace.get("cg_iters", 0) + n_hp
        trace.setdefault("cg_status", []).append(status)
    return w, status


def linesearch(g_model, H: AlHessian, w, w_l, w_u, fix_bounds):
    """src/basic_tralcnlss.jl:766-791."""
    wHw = H.vthv(w)
    with np.errstate(all="ignore"):
        alpha_opt = (-np.dot(g_model, w) / wHw) if wHw > 0 else np.inf
        alpha_allowed = np.inf
        free = ~fix_bounds
        neg = free & (w < 0)
        pos = free & (w > 0)
        if neg.any():
            alpha_allowed = min(alpha_allowed, float(np.min(w_l[neg] / w[neg])))
        if pos.any():
            alpha_allowed = min(alpha_allowed, float(np.min(w_u[pos] / w[pos])))
    return min(float(alpha_opt), alpha_allowed)


def minor_iterate(x, s, g_model, H: AlHessian, lincons: MixedConstraints, delta, kappa2, trace=None):
    """src/basic_tralcnlss.jl:649-675.  Trap T1: finite bounds land on the FIXED variables."""
    n = x.shape[0]
    x_minor = x + s
    w_u, w_l = np.full(n, np.inf), np.full(n, -np.inf)
    fx = lincons.fixvars
    w_u[fx] = np.minimum(lincons.xupp[fx] - x_minor[fx], delta)
    w_l[fx] = np.maximum(lincons.xlow[fx] - x_minor[fx], -delta)
    w, cg_status = projected_cg(g_model, H, w_l, w_u, lincons, kappa2, trace=trace)
    if cg_status != NEGATIVE_CURVATURE:
        alpha = linesearch(g_model, H, w, w_l, w_u, lincons.fixvars)
        w = alpha * w
    return w, cg_status


# --------------------------------------------------------------------------------------
# inner_step  (src/basic_tralcnlss.jl:394-460)
# --------------------------------------------------------------------------------------
def inner_step(x, g, H: AlHessian, chol_aat_L, lincons: MixedConstraints, delta, nb_minor_step, kappa2, kappa3,
               trace=None):
    """src/basic_tralcnlss.jl:394-460.  Mutates lincons.fixvars / lincons.chol."""
    m, n = lincons.lineq.shape
    s = cauchy_step(x, g, H, chol_aat_L, lincons, delta, trace=trace)  # :410
    g_minor = H.mul(s) + g  # :412
    j = 1
    norm_reduced_g = norm_reduced_gradient(g, lincons)
    norm_reduced_g_minor = norm_reduced_gradient(g_minor, lincons)
    approx_solved = norm_reduced_g_minor <= kappa3 * norm_reduced_g  # :423
    allowed_minor_step = n - m - lincons.nb_fix()  # :425 (1-arg max, trap T6)
    max_minor_step = min(nb_minor_step, allowed_minor_step)
    cg_stop = False
    n_minor = 0
    while j <= max_minor_step and (not approx_solved) and (not cg_stop):  # :430
        w, cg_status = minor_iterate(x, s, g_minor, H, lincons, delta, kappa2, trace=trace)
        cg_stop = cg_status == NEGATIVE_CURVATURE
        s = s + w  # :436
        g_minor = H.mul(s) + g  # :437
        active_indx = active_bounds(lincons, x, s, delta)  # :439
        if m + active_indx.shape[0] <= n:  # :441
            add_active(lincons, chol_aat_L, active_indx)
            norm_reduced_g = norm_reduced_gradient(g, lincons)
            norm_reduced_g_minor = norm_reduced_gradient(g_minor, lincons)
            approx_solved = norm_reduced_g_minor <= kappa3 * norm_reduced_g
        else:  # :450-452
            approx_solved = True
            active_bounds_reset(lincons, x + s, chol_aat_L)
        j += 1
        n_minor += 1
    model_reduction = float(np.dot(g, s) + 0.5 * H.vthv(s))  # :458
    if trace is not None:
        trace["minor_iters"] = trace.get("minor_iters", 0) + n_minor
    return s, model_reduction


# --------------------------------------------------------------------------------------
# Logging in the reference's benlsip.out format  (src/misc.jl:1-80)
# --------------------------------------------------------------------------------------
def _jl_e(prec: int, v: float) -> str:
    """Julia @sprintf("%.{p}e"): C formatting for finite values, `NaN` / `Inf` / `-Inf` otherwise (Julia's Printf)."""
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "Inf" if v > 0 else "-Inf"
    return f"%.{prec}e" % v


def print_tralcnllss_header(n, d, p, m, x_l, x_u, crit_tol, feas_tol, tau, eta1, eta2, gamma1, gamma2, io):
    """src/misc.jl:1-45 (argument names as in misc.jl; the caller swaps feas/crit, harmless)."""
    w = io.write
    w("\n\n")
    w("*" * 64 + "\n")
    w("*" + " " * 62 + "*\n")
    w("*" + " " * 23 + "BEnlsip.jl v-DEV" + " " * 23 + "*\n")
    w("*" + " " * 62 + "*\n")
    w("*                   Better version of ENLSIP                   *\n")
    w("*" + " " * 62 + "*\n")
    w("*" * 64 + "\n")
    w("\nProblem dimensions\n")
    w("Number of parameters.................: %5i\n" % n)
    w("Number of residuals..................: %5i\n" % d)
    w("Number of nonlinear constraints......: %5i\n" % p)
    w("Number of linear constraints.........: %5i\n" % m)
    w("Number of lower bounds...............: %5i\n" % int(np.count_nonzero(np.isfinite(x_l))))
    w("Number of upper bounds...............: %5i\n" % int(np.count_nonzero(np.isfinite(x_u))))
    w("\nAlgorithm parameters\n")
    w("Optimality tolerance.................................: %.6e\n" % crit_tol)
    w("Nonlinear constraints feasibility tolerance..........: %.6e\n" % feas_tol)
    w("Increase penalty parameter factor....................: %5f\n" % tau)
    w("Step acceptance treshold.............................: %5f\n" % eta1)
    w("Great step acceptance treshold.......................: %5f\n" % eta2)
    w("Trust region increase factor.........................: %5f\n" % gamma2)
    w("Trust region decrease factor.........................: %5f\n" % gamma1)
    w("\n\n\n")


def print_outer_iter_header(k, objective, nl_feas, mu, pix, omega, io, first=False):
    """src/misc.jl:47-68."""
    w = io.write
    w("\n" + "=" * 80 + "\n")
    w("                          Outer iter %d\n" % k)
    w("  objective    nl feasibility     μ      criticality   tolerance\n")
    if first:
        w("%s   %s  %s        -         %s" % (_jl_e(7, objective), _jl_e(6, nl_feas), _jl_e(2, mu), _jl_e(2, omega)))
    else:
        w("%s   %s  %s     %s     %s" % (_jl_e(7, objective), _jl_e(6, nl_feas), _jl_e(2, mu), _jl_e(2, pix), _jl_e(2, omega)))
    w("\n" + "=" * 80 + "\n")
    w("iter     AL value       ||s||        Δ          ρ\n")


def print_inner_iter(k, obj, norm_step, radius, rho, io):
    """src/misc.jl:70-80."""
    io.write("%4d   %s   %s   %s   %s\n" % (k, _jl_e(6, obj), _jl_e(2, norm_step), _jl_e(2, radius), _jl_e(2, rho)))


# --------------------------------------------------------------------------------------
# solve_subproblem  (src/basic_tralcnlss.jl:303-378)
# --------------------------------------------------------------------------------------
def solve_subproblem(x0, y, mu, residuals, nlconstraints, jac_res, jac_nlcons, chol_aat_L, lincons, nb_minor_step,
                     k_max, omega_tol, eta1, eta2, gamma1, gamma2, kappa2, kappa3, output_file=None, trace=None):
    """src/basic_tralcnlss.jl:303-378.  Returns (x, cx, pix)."""
    counters = None if trace is None else trace.setdefault("counters", {})
    x = np.array(x0, dtype=np.float64, copy=True)
    rx, cx, y_bar, mx, g, H = new_point(x0, y, mu, residuals, nlconstraints, jac_res, jac_nlcons, counters)
    pix = np.inf
    delta = initial_tr(g)
    k = 1
    solved = False
    while (not solved) and k <= k_max:
        s, pred = inner_step(x, g, H, chol_aat_L, lincons, delta, nb_minor_step, kappa2, kappa3, trace=trace)
        x_next = x + s
        rx_next, cx_next, mx_next = evaluate_al(x_next, y, mu, residuals, nlconstraints, counters)
        ared = mx_next - mx
        with np.errstate(all="ignore"):
            rho = float(np.float64(ared) / np.float64(pred))  # NaN when pred == 0 (trap T8)
        if output_file is not None:
            print_inner_iter(k, mx, float(np.linalg.norm(s)), delta, rho, output_file)
        if trace is not None:
            trace.setdefault("inner", []).append(
                dict(k=k, mx=mx, norm_s=float(np.linalg.norm(s)), delta=delta, rho=rho, pred=pred))
        if rho > eta1:
            x = x_next.copy()
            rx, cx, mx = rx_next.copy(), cx_next.copy(), mx_next
            y_bar, J, C, g = first_derivatives(x, y, mu, rx, cx, jac_res, jac_nlcons, counters)
            H = AlHessian(J, C, mu, counters)  # second_derivatives, :79-85
        delta = update_tr(delta, rho, eta1, eta2, gamma1, gamma2)
        pix = criticality_measure(g, lincons)
        if trace is not None:
            trace["inner"][-1]["pix"] = pix
            trace["inner"][-1]["nb_fix"] = lincons.nb_fix()
            trace["inner"][-1]["omega_tol"] = omega_tol
            trace["inner"][-1]["bp_cum"] = trace.get("breakpoints", 0)
            trace["inner"][-1]["cg_cum"] = trace.get("cg_iters", 0)
        solved = pix < omega_tol
        k += 1
    if trace is not None:
        trace["inner_iters"] = trace.get("inner_iters", 0) + (k - 1)
    return x, cx, pix


# --------------------------------------------------------------------------------------
# tralcnllss  (src/basic_tralcnlss.jl:167-298) -- the caller of the hot path
# --------------------------------------------------------------------------------------
def tralcnllss(x0, residuals, jac_res, nlconstraints, jac_nlcons, A, b, x_l, x_u, *,
               mu0=10.0, tau=100.0, omega0=1.0, eta0=1.0, feas_tol=SQRT_EPS, crit_tol=SQRT_EPS,
               k_crit=1.0, k_feas=0.1, beta_crit=1.0, beta_feas=0.9, eta1=0.25, eta2=0.75,
               gamma1=0.0625, gamma2=2.0, gamma_c=10.0, kappa1=1e-2, kappa2=0.1, kappa3=0.1,
               max_outer_iter=500, max_inner_iter=500, max_minor_iter=50, output_file=None, trace=None):
    """src/basic_tralcnlss.jl:167-298.  Returns (x, y).  `b` is never read (trap T9)."""
    assert (0 < eta1 <= eta2 < 1) and (0 < gamma1 < 1 < gamma2), "Invalid trust region updates paramaters"
    A = np.asarray(A, dtype=np.float64)
    m, n = A.shape
    chol_aat_L = _cholesky_lower(A @ A.T)  # :206
    x = np.array(x0, dtype=np.float64, copy=True)
    rx = residuals(x)
    cx = nlconstraints(x)
    mu = float(mu0)
    if output_file is not None:  # :213-226 (feas_tol / crit_tol swapped at the call site, as in the reference)
        print_tralcnllss_header(n, rx.shape[0], cx.shape[0], m, x_l, x_u, feas_tol, crit_tol, tau, eta1, eta2,
                                gamma1, gamma2, output_file)
    omega, eta = initial_tolerances(mu0, omega0, eta0, k_crit, k_feas)  # :229
    y = least_squares_multipliers(x, residuals, jac_res, jac_nlcons)  # :230
    polyhedron = MixedConstraints(A, chol_aat_L, l=x_l, u=x_u)  # :231
    first_order_critical = False
    outer_iter = 1
    if output_file is not None:
        print_outer_iter_header(outer_iter, float(np.dot(rx, rx)), float(np.linalg.norm(cx)), mu, 0.0, omega,
                                output_file, first=True)
    pix = np.inf
    while (not first_order_critical) and outer_iter <= max_outer_iter:  # :246
        x_next, cx_next, pix = solve_subproblem(x, y, mu, residuals, nlconstraints, jac_res, jac_nlcons, chol_aat_L,
                                                polyhedron, max_minor_iter, max_inner_iter, omega, eta1, eta2,
                                                gamma1, gamma2, kappa2, kappa3, output_file=output_file, trace=trace)
        feas_measure = float(np.linalg.norm(cx_next))
        if feas_measure <= eta:  # :273
            x[:] = x_next
            cx = np.array(cx_next, copy=True)
            first_order_critical = (pix <= crit_tol) and (feas_measure <= feas_tol)
            if not first_order_critical:
                y = first_order_multipliers(y, cx, mu)
                omega /= mu ** beta_crit
                eta /= mu ** beta_feas
        else:  # :284-289
            mu *= tau
            omega = omega0 / (mu ** k_crit)
            eta = eta0 / (mu ** k_feas)
        outer_iter += 1
        rx = residuals(x)
        objective = float(np.dot(rx, rx))  # :292
        if output_file is not None:
            print_outer_iter_header(outer_iter, objective, feas_measure, mu, pix, omega, output_file)
        if trace is not None:
            trace.setdefault("outer", []).append(
                dict(outer_iter=outer_iter, objective=objective, feas=feas_measure, mu=mu, pix=pix, omega=omega))
    if trace is not None:
        trace["outer_iters"] = outer_iter - 1
        trace["fixvars_words"] = polyhedron.fixvars_words()
        trace["x"] = x.copy()
        trace["y"] = np.array(y, copy=True)
        trace["mu"] = mu
    return x, y


def is_feasible(x, A, x_l, x_u, b):
    """src/basic_tralcnlss.jl:142-150 (`isapprox` default rtol = sqrt(eps))."""
    Ax = A @ x
    ok = np.linalg.norm(Ax - b) <= SQRT_EPS * max(np.linalg.norm(Ax), np.linalg.norm(b))
    return bool(ok and np.all(x_l <= x) and np.all(x <= x_u))
