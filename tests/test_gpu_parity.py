"""
GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the oracle
(oracle/benlsip_oracle.py) on the same seeded inputs.  Floating point: FP64, tolerances written at each assert
(north_star: final iterate / objective within 1e-10 relative, active-set words bit-exact, iteration counts equal).
"""
import numpy as np
import pytest

import benlsip_b200 as B
from oracle import benlsip_oracle as O
from oracle.models import DenseExpSumProblem, ExpSumProblem, GlmProblem, MixedConstraintProblem, SphereRegression

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def _golden(name):
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".json")))


def assert_matches_golden(name, tr_g, x_g, tol=1e-10):
    """Against the committed golden (generated once by the oracle, tests/golden/make_golden.py): exact up to the golden's first
    noise-driven decision, everything exact when it has none (tests/parity.py states the criterion)."""
    from tests.parity import assert_trajectory_parity
    return assert_trajectory_parity(name, tr_g, x_g, tol=tol)


def assert_parity_with_live_oracle(name, tr_g, x_g, tr_o, x_o, tol=1e-10):
    """Against the oracle run on THIS host.  The exact assertions are the golden ones (assert_matches_golden, always
    enforced).  NumPy/OpenBLAS sums in a host-dependent order (kernel, thread count), so the ORACLE's own late iterations
    can deviate from its committed golden on some hosts: that case is reported as an explicit xfail of the live comparison
    (it says nothing about the CUDA path, which has already matched the golden); otherwise the comparison is tight."""
    from tests.parity import first_fragile
    g = _golden(name)
    F = first_fragile(g)
    nprefix = len(g["inner"]) if F is None else F
    oracle_counts = (tr_o["outer_iters"], tr_o["inner_iters"], tr_o.get("cg_iters", 0), tr_o.get("breakpoints", 0))
    golden_counts = (g["outer_iters"], g["inner_iters"], g["cg_iters"], g["breakpoints"])
    if F is None and oracle_counts != golden_counts:
        pytest.xfail(f"live oracle on this host deviates from its own committed golden {name}: {oracle_counts} vs {golden_counts} "
                     "(host BLAS summation order); the CUDA path matched the golden exactly")
    st = tr_g["stats"]
    for a, b in list(zip(tr_g["inner"], tr_o["inner"]))[:nprefix]:
        assert (a["k"], a["nb_fix"], a["bp_cum"], a["cg_cum"]) == (b["k"], b["nb_fix"], b["bp_cum"], b["cg_cum"])
        assert abs(a["mx"] - b["mx"]) <= 1e-10 * abs(b["mx"])
    if F is None:
        assert (tr_g["outer_iters"], st["inner_iters"], st["cg_iters"], st["breakpoints"]) == oracle_counts
        assert rel(x_g, x_o) < tol
        assert np.array_equal(tr_g["fixvars_words"], tr_o["fixvars_words"])
    else:
        assert rel(x_g, x_o) < 2e-8


@pytest.fixture()
def S():
    s = B.Solver(0)
    yield s
    s.close()


# ---------------------------------------------------------------------------------------------------------
# K1-K4: the streaming Jacobian kernels, every tiling regime (TG/G/KCH/RB), ragged M, odd n
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,n", [(4, 3), (1, 1), (33, 5), (1000, 37), (4096, 256), (3001, 250), (2500, 1024), (777, 1000),
                                 (700, 2048), (300, 4096), (129, 3000), (64, 16), (5000, 64), (200000, 128), (90, 8192), (75, 5000)])
def test_alhessian_matvecs_match_numpy(S, M, n):
    """test/structures.jl:1-16 generalised: H*v, vthv, J*v, J'*w against explicit NumPy (rtol 1e-12)."""
    rng = np.random.default_rng(M * 7919 + n)
    J = rng.standard_normal((M, n))
    v = rng.standard_normal(n)
    w = rng.standard_normal(M)
    S.set_problem(M, n)
    S.upload_jacobian(J)
    Jv = J @ v
    assert rel(S.hess_mul(v), J.T @ Jv) < 1e-12
    assert abs(S.vthv(v) - Jv @ Jv) <= 1e-12 * (Jv @ Jv)
    assert rel(S.jv(v), Jv) < 1e-12
    assert rel(S.jtw(w), J.T @ w) < 1e-12


def test_hess_mul_is_run_to_run_deterministic(S):
    rng = np.random.default_rng(5)
    J = rng.standard_normal((20000, 512))
    v = rng.standard_normal(512)
    S.set_problem(20000, 512)
    S.upload_jacobian(J)
    a = S.hess_mul(v)
    for _ in range(3):
        assert np.array_equal(a, S.hess_mul(v))  # fixed reduction order, no atomics


def test_alhessian_with_nonlinear_constraint_block(S):
    """test/structures.jl:1-16 verbatim shape: H = J'J + mu C'C, n = 5."""
    rng = np.random.default_rng(0)
    n = 5
    J, Cm, mu, v = rng.random((n, n)), rng.random((n, n)), rng.random(), rng.random(n)
    S.set_problem(n, n, p=n)
    S.upload_jacobian(J)
    S.upload_nlcons_jacobian(Cm)
    S.set_mu(mu)
    H_test = J.T @ J + mu * Cm.T @ Cm
    assert rel(S.hess_mul(v), H_test @ v) < 1e-13
    assert abs(S.vthv(v) - v @ (H_test @ v)) < 1e-13 * abs(v @ (H_test @ v))


def test_gram_dmma_matches_numpy(S):
    rng = np.random.default_rng(11)
    for M, n in [(3000, 200), (5000, 384), (1000, 130)]:
        J = rng.standard_normal((M, n))
        S.set_problem(M, n)
        S.upload_jacobian(J)
        G, ms = S.gram()
        ref = J.T @ J
        assert np.max(np.abs(G - ref)) <= 1e-12 * np.max(np.abs(ref))
        assert np.array_equal(G, G.T)


# ---------------------------------------------------------------------------------------------------------
# K11: device models against oracle/models.py
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,Mtot,n,nranks,rank", [("glm", 4096, 64, 1, 0), ("glm", 3001, 250, 1, 0), ("glm", 71500, 1024, 8, 7),
                                                     ("glm", 9000, 96, 2, 1), ("expsum", 4096, 16, 1, 0), ("expsum", 5140, 64, 4, 2),
                                                     ("expsum_dense", 4096, 16, 1, 0), ("expsum_dense", 6000, 256, 2, 1)])
def test_builtin_models_match_oracle(S, kind, Mtot, n, nranks, rank):
    """Rows [row0, row0 + M) of a global problem: the shard of `rank` of `nranks` (bnl_shard_rows)."""
    row0, M = B.shard_rows(Mtot, nranks, rank)
    if kind == "glm":
        P = GlmProblem(M, n, seed=3, row0=row0)
        mid = B.MODEL_GLM
        seed = 3
    elif kind == "expsum":
        P = ExpSumProblem(M, n, seed=1, row0=row0, M_total=Mtot)
        mid = B.MODEL_EXPSUM
        seed = 1
    else:
        P = DenseExpSumProblem(M, n, seed=1, row0=row0, M_total=Mtot)
        mid = B.MODEL_EXPSUM_DENSE
        seed = 1
    S.set_problem(M, n, M_total=Mtot, row0=row0)
    S.use_builtin_model(mid, noise=1e-3, cond_exp=0.0, seed=seed)
    mv = S.model_vectors()
    assert np.array_equal(mv["x_true"], P.x_true)  # same hash, exact
    assert np.array_equal(mv["xlow"], P.xlow) and np.array_equal(mv["xupp"], P.xupp) and np.array_equal(mv["x0"], P.x0)
    x = P.x0 + 0.05 * np.cos(np.arange(n))
    r, ss = S.residuals(x)
    r_ref = P.residuals(x)
    assert np.max(np.abs(r - r_ref)) <= 1e-13 * max(1.0, np.max(np.abs(r_ref)))
    assert abs(ss - r_ref @ r_ref) <= 1e-12 * (r_ref @ r_ref)
    S.eval_jacobian(x)
    J = P.jac_res(x)
    v = np.sin(np.arange(n) + 1.0)
    assert rel(S.jv(v), J @ v) < 1e-12
    assert rel(S.hess_mul(v), J.T @ (J @ v)) < 1e-12
    for k in (0, n // 2, n - 1):  # individual columns of J
        e = np.zeros(n)
        e[k] = 1.0
        assert np.max(np.abs(S.jv(e) - J[:, k])) <= (1e-14 if kind != "expsum_dense" else 1e-13) * max(1.0, np.max(np.abs(J[:, k])))


# ---------------------------------------------------------------------------------------------------------
# K8-K10: MixedConstraints through the library (reference fixtures, test/structures.jl:18-78)
# ---------------------------------------------------------------------------------------------------------
def test_hs48_projection_golden_vector(S):
    """test/structures.jl:37-58: the reference's only literal golden vector, through the CUDA path."""
    A = np.array([[1.0, 1, 1, 1, 1], [0, 0, 1, -2, -2]])
    x_hs = np.array([3.0, 5, -3, 2, -2])
    S.set_problem(1, 5, A)
    S.set_fixvars([True, True, False, False, False])
    Bm = np.vstack([A, np.eye(5)[[0, 1], :]])
    yv = np.array([0.3, -0.7, 1.1, 0.25])
    np.testing.assert_allclose(S.left_mul_tr(yv), Bm.T @ yv, rtol=1e-14)  # test/structures.jl:49
    np.testing.assert_allclose(S.left_mul(x_hs), Bm @ x_hs, rtol=1e-14)   # test/structures.jl:50
    proj = S.projection(x_hs)
    np.testing.assert_allclose(proj, [0.0, 0, 0, 2, -2], rtol=0, atol=1e-14)
    v = A @ proj
    # reference asserts <= eps() with LAPACK's rounding (test/structures.jl:56); our factor differs in the last ulps
    assert np.all(np.abs(proj[:2]) <= 8 * np.finfo(float).eps) and v @ v <= 8 * np.finfo(float).eps
    # and against the oracle on a random vector
    cons = O.MixedConstraints(A, np.linalg.cholesky(A @ A.T), fixed=np.array([True, True, False, False, False]))
    r = np.array([0.3, -1.2, 2.5, 0.7, -0.1])
    np.testing.assert_allclose(S.projection(r), O.projection(cons, r), rtol=0, atol=1e-14)


def test_block_cholesky_equals_greedy(S):
    """test/structures.jl:18-35."""
    rng = np.random.default_rng(1)
    m, n = 3, 6
    A = rng.random((m, n))
    S.set_problem(1, n, A, -rng.random(n), rng.random(n) + 1)
    act = [1, 3, 5]
    fixed = np.zeros(n, dtype=bool)
    fixed[act] = True
    S.set_fixvars(fixed)
    Bm = np.vstack([A, np.eye(n)[act, :]])
    np.testing.assert_allclose(S.chol_L(), np.linalg.cholesky(Bm @ Bm.T), rtol=1e-10, atol=1e-12)
    assert S.fixvars().tolist() == fixed.tolist()


def test_active_bounds_identification_and_update(S):
    """test/structures.jl:60-78."""
    rng = np.random.default_rng(3)
    m, n = 3, 7
    A = rng.random((m, n))
    S.set_problem(1, n, A, -10 * np.ones(n), 10 * np.ones(n))
    x = rng.random(n)
    x[1] = -10.0
    S.active_bounds_reset(x)
    assert S.fixvars().tolist() == [False, True, False, False, False, False, False]
    S.add_active([2, 4])
    S.add_active(6)
    assert S.fixvars().tolist() == [False, True, True, False, True, False, True]
    assert S.fixvars_words().tolist() == [0b1010110]
    # active_bounds (polyhedral_constraints.jl:219-237) against the oracle, trust-region faces included
    cons = O.MixedConstraints(A, np.linalg.cholesky(A @ A.T), l=-10 * np.ones(n), u=10 * np.ones(n))
    s = np.array([0.5, 0.0, -0.5, 0.2, 0.5, -0.1, 0.3])
    assert S.active_bounds(x, s, 0.5).tolist() == O.active_bounds(cons, x, s, 0.5).tolist()


def test_mask_projection_is_exact_for_bound_only(S):
    n = 70
    S.set_problem(1, n, None, -np.ones(n), np.ones(n))
    r = np.random.default_rng(2).standard_normal(n)
    assert np.array_equal(S.projection(r), r)
    S.add_active([1, 3, 64, 69])
    v = S.projection(r)
    fix = np.zeros(n, dtype=bool)
    fix[[1, 3, 64, 69]] = True
    assert np.array_equal(v[~fix], r[~fix]) and not v[fix].any()
    assert S.fixvars_words().tolist() == [(1 << 1) | (1 << 3), (1 << 0) | (1 << 5)]
    with pytest.raises(IndexError):
        S.add_active(n)


def test_error_mapping(S):
    with pytest.raises(AssertionError):  # src/basic_tralcnlss.jl:200
        S.set_params(eta1=0.9, eta2=0.5)
    with pytest.raises(B.PosDefException):  # cholesky(A*A') of a rank-deficient A, :206
        S.set_problem(1, 4, np.array([[1.0, 2, 3, 4], [2.0, 4, 6, 8]]))


# ---------------------------------------------------------------------------------------------------------
# Step computation against the oracle: cauchy_step, projected_cg, inner_step
# ---------------------------------------------------------------------------------------------------------
def _glm_state(S, M, n):
    P = GlmProblem(M, n, seed=3)
    S.set_problem(M, n)
    S.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, 3)
    x = P.x0.copy()
    x[::7] = 1.0  # some variables start on their upper bound
    x[3::11] = -1.0
    S.eval_jacobian(x)
    J, r = P.jac_res(x), P.residuals(x)
    g = J.T @ r
    H = O.AlHessian(J, np.zeros((0, n)), 0.0)
    L0 = O._cholesky_lower(np.zeros((0, 0)))
    cons = O.MixedConstraints(P.A, L0, l=P.xlow, u=P.xupp)
    return P, x, g, H, L0, cons


@pytest.mark.parametrize("delta", [1e-3, 0.05, 10.0])
def test_cauchy_step_matches_oracle(S, delta):
    P, x, g, H, L0, cons = _glm_state(S, 3000, 96)
    s_ref = O.cauchy_step(x, g, H, L0, cons, delta)
    s = S.cauchy_step(x, g, delta)
    assert rel(s, s_ref) < 1e-12
    assert np.array_equal(S.fixvars_words(), cons.fixvars_words())


def test_inner_step_matches_oracle(S):
    P, x, g, H, L0, cons = _glm_state(S, 3000, 96)
    for delta in (0.02, 0.5):
        tr = {}
        s_ref, pred_ref = O.inner_step(x, g, H, L0, cons, delta, 50, 0.1, 0.1, trace=tr)
        S.reset_stats()
        s, pred = S.inner_step(x, g, delta)
        st = S.stats()
        assert rel(s, s_ref) < 1e-10
        assert abs(pred - pred_ref) <= 1e-10 * abs(pred_ref)
        assert np.array_equal(S.fixvars_words(), cons.fixvars_words())
        assert st["cg_iters"] == tr.get("cg_iters", 0) and st["minor_iters"] == tr.get("minor_iters", 0)
        assert st["breakpoints"] == tr.get("breakpoints", 0)


def test_projected_cg_matches_oracle(S):
    P, x, g, H, L0, cons = _glm_state(S, 3000, 96)
    O.active_bounds_reset(cons, x, L0)
    S.active_bounds_reset(x)
    s = np.zeros(96)
    g_minor = g.copy()
    n = 96
    w_u, w_l = np.full(n, np.inf), np.full(n, -np.inf)
    fx = cons.fixvars
    w_u[fx] = np.minimum(cons.xupp[fx] - x[fx], 1.0)
    w_l[fx] = np.maximum(cons.xlow[fx] - x[fx], -1.0)
    w_ref, st_ref = O.projected_cg(g_minor, H, w_l, w_u, cons, 0.1)
    w, st, iters = S.projected_cg(x, s, g_minor, 1.0)
    assert st == st_ref
    assert rel(w, w_ref) < 1e-11


# ---------------------------------------------------------------------------------------------------------
# Full solves: outer loop on the host (Python standing in for Julia), subproblems through the C ABI
# ---------------------------------------------------------------------------------------------------------
def _solve_both(S, P, model_id, seed, **kw):
    tr_o, tr_g = {}, {}
    x_o, y_o = O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp,
                            trace=tr_o, **kw)
    S.set_problem(P.M, P.n)
    S.use_builtin_model(model_id, 1e-3, 0.0, seed)
    x_g, y_g = B.tralcnllss(P.x0, None, None, None, None, None, None, None, None, solver=S, trace=tr_g, **kw)
    return x_o, tr_o, x_g, tr_g


@pytest.mark.parametrize("M,n", [(4096, 64), (20000, 256), (6000, 1024)])
def test_glm_full_solve_parity(S, M, n):
    """cfg3 family, shrunk: same outer/inner/minor/CG/breakpoint counts, x within 1e-10, active-set words bit-exact, per-inner-
    iteration AL values -- against the committed golden (exact) and against the oracle run live on this host."""
    P = GlmProblem(M, n, seed=3)
    x_o, tr_o, x_g, tr_g = _solve_both(S, P, B.MODEL_GLM, 3)
    assert_matches_golden(f"glm_{M}_{n}", tr_g, x_g)
    obj_g, obj_o = S.residuals(x_g, False)[1], _golden(f"glm_{M}_{n}")["objective"]
    assert abs(obj_g - obj_o) <= 1e-10 * obj_o
    # the literal search (a Hessian apply per breakpoint, :633) gives the SAME iterate bit for bit: the default device-side
    # breakpoint loop only ever lets literal numbers reach the iterate
    st_inc = tr_g["stats"]
    assert st_inc["inc_breakpoints"] == st_inc["breakpoints"] and st_inc["cauchy_loop_launches"] >= st_inc["inner_iters"]
    S.set_cauchy_mode(B.CAUCHY_LITERAL)
    S.reset_stats()
    tr_l = {}
    x_l, _ = B.tralcnllss(P.x0, None, None, None, None, None, None, None, None, solver=S, trace=tr_l)
    S.set_cauchy_mode(B.CAUCHY_INCREMENTAL)
    assert np.array_equal(x_l, x_g) and np.array_equal(tr_l["fixvars_words"], tr_g["fixvars_words"])
    assert tr_l["stats"]["inc_breakpoints"] == 0 and tr_l["stats"]["breakpoints"] == st_inc["breakpoints"]
    assert tr_l["stats"]["j_passes"] > st_inc["j_passes"]
    assert_matches_golden(f"glm_{M}_{n}", tr_l, x_l)
    assert_parity_with_live_oracle(f"glm_{M}_{n}", tr_g, x_g, tr_o, x_o)


def _assert_trace_prefix(tr_g, tr_o, nprefix, mx_rtol=1e-12, pix_rtol=1e-5):
    assert len(tr_g["inner"]) >= nprefix and len(tr_o["inner"]) >= nprefix
    for a, b in list(zip(tr_g["inner"], tr_o["inner"]))[:nprefix]:
        assert a["k"] == b["k"] and a["nb_fix"] == b["nb_fix"]
        assert abs(a["mx"] - b["mx"]) <= mx_rtol * abs(b["mx"])
        assert abs(a["delta"] - b["delta"]) <= 1e-9 * abs(b["delta"])
        assert abs(a["pix"] - b["pix"]) <= pix_rtol * abs(b["pix"])


def test_expsum_full_solve_parity(S):
    """cfg2 family, shrunk.  On this family the reference algorithm creeps (hundreds of inner iterations whose AL
    decrease is ~1e-14 relative), so rho = ared/pred is a ratio of rounding noise and the iteration at which
    pix < omega is crossed is not reproducible between two correct FP64 implementations (measured: 188 vs 201
    inner iterations, DESIGN.md 'parity floor').  What is reproducible and asserted: the first 60 inner iterations
    (AL value to 1e-12, criticality to 1e-5, active-set size, k), the outer count, the final iterate to 1e-8 and the
    final active set bit-exactly."""
    P = ExpSumProblem(4096, 16, seed=1)
    x_o, tr_o, x_g, tr_g = _solve_both(S, P, B.MODEL_EXPSUM, 1)
    assert tr_g["outer_iters"] == tr_o["outer_iters"]
    _assert_trace_prefix(tr_g, tr_o, 60)
    assert abs(tr_g["stats"]["inner_iters"] - tr_o["inner_iters"]) <= 0.15 * tr_o["inner_iters"]
    assert np.max(np.abs(x_g - x_o)) < 1e-8
    assert np.array_equal(tr_g["fixvars_words"], tr_o["fixvars_words"])


def test_dense_expsum_first_inner_steps_match_oracle(S):
    """cfg2 as SURVEY 8d words it (ONE n/2-term exponential sum on one time grid): kappa(J) ~ 1e16 at x0, the reference algorithm
    does not converge on it (oracle: max_inner_iter exhausted in every outer iteration), so parity is asserted where it is
    well-posed: model values, the Cauchy step (no CG involved) and the first inner step's counts."""
    M, n = 4096, 16
    P = DenseExpSumProblem(M, n, seed=1)
    S.set_problem(M, n)
    S.use_builtin_model(B.MODEL_EXPSUM_DENSE, 1e-3, 0.0, 1)
    x = P.x0.copy()
    mx, g, _ = S.new_point(x, None, 10.0)
    J, r = P.jac_res(x), P.residuals(x)
    assert abs(mx - 0.5 * r @ r) <= 1e-12 * (0.5 * r @ r) and rel(g, J.T @ r) < 1e-11
    H = O.AlHessian(J, np.zeros((0, n)), 0.0)
    L0 = O._cholesky_lower(np.zeros((0, 0)))
    for delta in (1e-3, 0.1):
        cons = O.MixedConstraints(P.A, L0, l=P.xlow, u=P.xupp)
        s_ref = O.cauchy_step(x, J.T @ r, H, L0, cons, delta)
        s = S.cauchy_step(x, J.T @ r, delta)
        assert rel(s, s_ref) < 1e-9 and np.array_equal(S.fixvars_words(), cons.fixvars_words())


def test_sphere_regression_through_callbacks():
    """cfg1: test/problems/sphere_regression.jl end to end through the library (m_lin = 1, p = 1, callbacks, general
    projection with device Cholesky / triangular solves), with the reference's end-state assertions
    (:63-65).  Trajectory: a 1-ulp perturbation of the residuals already changes the ORACLE's own counts on this
    problem (8 -> 7 outer, 57 -> 56 inner, x by 5e-9, y by 7e-8: tests/test_oracle_noise_floor.py), so parity is
    asserted on the first 10 inner iterations and on the end state to that floor."""
    P = SphereRegression
    tr_o, tr_g = {}, {}
    x_o, y_o = O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp,
                            max_outer_iter=100, max_inner_iter=250, trace=tr_o)
    x_g, y_g = B.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp,
                            max_outer_iter=100, max_inner_iter=250, trace=tr_g)
    assert np.linalg.norm(P.nlconstraints(x_g)) < O.SQRT_EPS  # :63
    assert O.is_feasible(x_g, P.A, P.xlow, P.xupp, P.b)  # :64
    grad_lag = P.jac_res(x_g).T @ P.residuals(x_g) + P.jac_nlcons(x_g).T @ y_g
    from tests.test_oracle_reference_fixtures import _project_polyhedron_small
    # :65 asserts < 1e-7; under +-1 ulp perturbations of r the ORACLE itself lands at 0.7e-7 .. 4e-7 (y lags one
    # multiplier update when the loop exits, :276-283), so the reproducible bound is 1e-6
    assert np.linalg.norm(x_g - _project_polyhedron_small(x_g - grad_lag, P.A, P.b, P.xlow, P.xupp)) < 1e-6
    _assert_trace_prefix(tr_g, tr_o, 10, mx_rtol=1e-12, pix_rtol=1e-6)
    assert abs(tr_g["outer_iters"] - tr_o["outer_iters"]) <= 1
    assert np.max(np.abs(x_g - x_o)) < 5e-8 and np.max(np.abs(y_g - y_o)) < 5e-7
    # (the FINAL active set belongs to the noise-driven tail -- trust-region faces of a radius that rounding noise decides,
    # tests/parity.py -- so it is compared through nb_fix on the well-conditioned prefix above, not at the end)


def test_cfg4_family_mixed_constraints_full_solve_parity():
    """cfg4 family (shrunk, host callbacks): m_lin = 4 linear equalities + 1 nonlinear sphere constraint + box, the AL
    loop really exercised (mu goes 10 -> 1e9).  General projection on the device (Cholesky of A~A~', rebuilt on every
    add_active!, triangular solves).  The oracle is insensitive to 1-ulp perturbations on this problem (same counts,
    dx 9e-16), so counts are asserted exactly."""
    P = MixedConstraintProblem(600, 24, 4)
    kw = dict(max_outer_iter=60, max_inner_iter=200)
    tr_o, tr_g = {}, {}
    x_o, y_o = O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, trace=tr_o, **kw)
    x_g, y_g = B.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, trace=tr_g, **kw)
    st = tr_g["stats"]
    assert tr_g["mu"] == _golden("mixed_600_24_4")["mu"] and tr_g["mu"] > 10.0
    assert_matches_golden("mixed_600_24_4", tr_g, x_g)
    assert rel(y_g, np.array(_golden("mixed_600_24_4")["y"])) < 1e-8
    assert abs(P.nlconstraints(x_g)[0]) < 1e-8 and np.max(np.abs(P.A @ x_g - P.b)) < 1e-12
    assert st["chol_rebuilds"] > 0
    assert_parity_with_live_oracle("mixed_600_24_4", tr_g, x_g, tr_o, x_o)


def test_cfg4_family_on_device_matches_oracle(S):
    """cfg4 family with NO host callbacks: device GLM residual/Jacobian generators (truth vector supplied), built-in sphere
    constraint (bnl_use_builtin_nlcons), linear equalities A from the host once (bnl_set_problem), general projection on
    the device.  Same problem as MixedConstraintProblem => exact counts against the oracle."""
    P = MixedConstraintProblem(600, 24, 4)
    S.set_problem(P.M, P.n, P.A, P.xlow, P.xupp, p=1)
    S.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, P.seed)
    S.model_set_truth(P.x_star, P.x0)
    S.use_builtin_nlcons(B.NLCONS_SPHERE, P.rho2)
    r, ss = S.residuals(P.x0)
    assert np.max(np.abs(r - P.residuals(P.x0))) < 1e-13
    c, Cm = S.nlcons(P.x0)
    assert abs(c[0] - P.nlconstraints(P.x0)[0]) < 1e-13 and np.allclose(Cm, P.jac_nlcons(P.x0), rtol=0, atol=0)
    assert rel(S.gradient(P.x0), P.jac_res(P.x0).T @ P.residuals(P.x0)) < 1e-12
    kw = dict(max_outer_iter=60, max_inner_iter=200)
    tr_o, tr_g = {}, {}
    x_o, y_o = O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, trace=tr_o, **kw)
    x_g, y_g = B.tralcnllss(P.x0, None, None, None, None, None, None, None, None, solver=S, trace=tr_g, **kw)
    assert_matches_golden("mixed_600_24_4", tr_g, x_g)
    assert tr_g["mu"] == _golden("mixed_600_24_4")["mu"] and rel(y_g, np.array(_golden("mixed_600_24_4")["y"])) < 1e-8
    assert_parity_with_live_oracle("mixed_600_24_4", tr_g, x_g, tr_o, x_o)


def test_cfg5_family_ill_conditioned_inner_steps(S):
    """cfg5 family (column scaling 10^(-6 j/n), kappa(J'J) ~ 1e12), shrunk.  The reference algorithm does not reach its
    tolerance on this family in any reasonable number of iterations (oracle: 20 outer x 100 inner x ~50 CG at n = 64), so
    parity is asserted per inner step against the oracle's committed steps (tests/golden/glm_cfg5_3000_96_steps.json):
    Cauchy point + projected CG iterates, predicted reduction, CG / breakpoint counts, active set."""
    G5 = _golden("glm_cfg5_3000_96_steps")
    M, n = G5["M"], G5["n"]
    P = GlmProblem(M, n, seed=3, cond_exp=G5["cond_exp"])
    S.set_problem(M, n)
    S.use_builtin_model(B.MODEL_GLM, 1e-3, G5["cond_exp"], 3)
    assert np.allclose(S.model_vectors()["x_true"], P.x_true)
    for stp in G5["steps"]:
        x, g = np.array(stp["x"]), np.array(stp["g"])
        S.eval_jacobian(x)
        assert rel(S.jtw(S.residuals(x)[0]), g) < 1e-11
        S.reset_stats()
        s, pred = S.inner_step(x, g, stp["delta"])
        st = S.stats()
        assert st["cg_iters"] == stp["cg_iters"] and st["breakpoints"] == stp["breakpoints"] and st["minor_iters"] == stp["minor_iters"]
        assert rel(s, np.array(stp["s"])) < 1e-7  # CG on kappa ~ 1e12 amplifies rounding; the step still agrees to 7 digits
        assert abs(pred - stp["pred"]) <= 1e-8 * abs(stp["pred"])
        assert [int(w) for w in S.fixvars_words()] == stp["fixvars_words"]


@pytest.mark.parametrize("M,n", [(4096, 64), (20000, 256)])
def test_gram_mode_solve_matches_matrix_free_and_oracle(S, M, n):
    """Opt-in Gram-apply mode (G = J'J on the FP64 tensor cores once per Jacobian).  H*v from G has different
    rounding than J'(Jv), so the last outer iteration may flicker (measured: 8 vs 7 outer at (4096,64), identical counts at
    (20000,256)); asserted: final iterate within 5e-9 of the oracle, outer count within 1, active set bit-exact."""
    P = GlmProblem(M, n, seed=3)
    tr_o, tr_g = {}, {}
    x_o, _ = O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, trace=tr_o)
    S.set_problem(P.M, P.n)
    S.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, 3)
    S.set_hessian_mode(B.HESSIAN_GRAM)
    x_g, _ = B.tralcnllss(P.x0, None, None, None, None, None, None, None, None, solver=S, trace=tr_g)
    st = tr_g["stats"]
    assert st["gram_count"] == st["jac_eval"] and st["j_passes"] < st["hess_mul"]
    assert abs(tr_g["outer_iters"] - tr_o["outer_iters"]) <= 1 and abs(st["inner_iters"] - tr_o["inner_iters"]) <= 2
    assert rel(x_g, x_o) < 5e-9  # measured 1.3e-9 at (4096,64) (one extra outer iteration), < 1e-10 at (20000,256)
    assert np.array_equal(tr_g["fixvars_words"], tr_o["fixvars_words"])
    # H*v from G against the matrix-free kernel
    v = np.cos(np.arange(n) * 0.37)
    hv_g, q_g = S.hess_mul(v), S.vthv(v)
    S.set_hessian_mode(B.HESSIAN_MATRIX_FREE)
    hv_f, q_f = S.hess_mul(v), S.vthv(v)
    assert rel(hv_g, hv_f) < 1e-12 and abs(q_g - q_f) <= 1e-12 * q_f


def test_device_cauchy_loop_equals_literal_search(S):
    """cauchy_step / inner_step through the default device-side breakpoint loop vs the literal search (:574-639): the Cauchy
    point, the step, the predicted reduction and the active set are bit-identical; only the pass count differs."""
    P, x, g, H, L0, cons = _glm_state(S, 3000, 96)
    for delta in (1e-3, 0.05, 10.0):
        S.set_cauchy_mode(B.CAUCHY_LITERAL)
        S.reset_stats()
        s_lit, pred_lit = S.inner_step(x, g, delta)
        w_lit, st_lit = S.fixvars_words(), S.stats()
        S.set_cauchy_mode(B.CAUCHY_INCREMENTAL)
        S.reset_stats()
        s_inc, pred_inc = S.inner_step(x, g, delta)
        st_inc = S.stats()
        assert np.array_equal(s_inc, s_lit) and pred_inc == pred_lit
        assert np.array_equal(S.fixvars_words(), w_lit)
        assert st_inc["breakpoints"] == st_lit["breakpoints"] == st_inc["inc_breakpoints"]
        assert st_inc["cauchy_loop_launches"] >= 1 and st_inc["j_passes"] <= st_lit["j_passes"]


def test_device_cauchy_loop_guard_band_forces_literal_evaluations(S):
    """With an absurdly wide rounding band every decision is 'inside the band': the loop must hand EVERY breakpoint to the
    literal evaluation and still produce the literal result (the guard path is the one that protects parity)."""
    import os
    P, x, g, H, L0, cons = _glm_state(S, 3000, 96)
    S.set_cauchy_mode(B.CAUCHY_LITERAL)
    s_lit, pred_lit = S.inner_step(x, g, 0.05)
    w_lit = S.fixvars_words()
    os.environ["BNL_CAUCHY_GUARD"] = "1e30"
    try:
        S2 = B.Solver(0)
    finally:
        del os.environ["BNL_CAUCHY_GUARD"]
    S2.set_problem(3000, 96)
    S2.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, 3)
    S2.eval_jacobian(x)
    S2.reset_stats()
    s2, pred2 = S2.inner_step(x, g, 0.05)
    st = S2.stats()
    assert np.array_equal(s2, s_lit) and pred2 == pred_lit and np.array_equal(S2.fixvars_words(), w_lit)
    assert st["cauchy_literal_evals"] >= st["breakpoints"] >= 1
    S2.close()
    S.set_cauchy_mode(B.CAUCHY_INCREMENTAL)


def test_device_cauchy_loop_with_nonlinear_constraint_block(S):
    """Bounds + one nonlinear (sphere) constraint, no linear equalities: the loop carries the mu C'C part of AlHessian
    (:92-106) itself; result bit-identical to the literal search, which applies H = J'J + mu C'C per breakpoint."""
    M, n = 2000, 48
    P = GlmProblem(M, n, seed=3)
    S.set_problem(M, n, None, P.xlow, P.xupp, p=1)
    S.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, 3)
    S.use_builtin_nlcons(B.NLCONS_SPHERE, 3.0)
    x = P.x0.copy()
    x[::5] = 0.4
    y = np.array([0.3])
    for mu in (10.0, 1e4):
        res = []
        for mode in (B.CAUCHY_LITERAL, B.CAUCHY_INCREMENTAL):
            S.set_cauchy_mode(mode)
            mx, g, cx = S.new_point(x, y, mu)
            S.reset_stats()
            s, pred = S.inner_step(x, g, 0.3)
            res.append((s, pred, S.fixvars_words(), S.stats()))
        assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1] and np.array_equal(res[0][2], res[1][2])
        assert res[1][3]["inc_breakpoints"] == res[1][3]["breakpoints"] == res[0][3]["breakpoints"]
        # and against the oracle with the same H
        J, r = P.jac_res(x), P.residuals(x)
        Cm = (2.0 * x)[None, :]
        c = np.array([x @ x - 3.0])
        g_ref = J.T @ r + Cm.T @ (y + mu * c)
        assert rel(g, g_ref) < 1e-12
        H = O.AlHessian(J, Cm, mu)
        L0 = O._cholesky_lower(np.zeros((0, 0)))
        cons = O.MixedConstraints(np.zeros((0, n)), L0, l=P.xlow, u=P.xupp)
        s_ref, pred_ref = O.inner_step(x, g_ref, H, L0, cons, 0.3, 50, 0.1, 0.1)
        assert rel(res[1][0], s_ref) < 1e-9 and abs(res[1][1] - pred_ref) <= 1e-9 * abs(pred_ref)
        assert np.array_equal(res[1][2], cons.fixvars_words())
    S.set_cauchy_mode(B.CAUCHY_INCREMENTAL)


def test_native_outer_loop_equals_host_outer_loop(S, tmp_path):
    P = GlmProblem(4096, 64, seed=3)
    S.set_problem(P.M, P.n)
    S.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, 3)
    x_h, _ = B.tralcnllss(P.x0, None, None, None, None, None, None, None, None, solver=S)
    w_h = S.fixvars_words()
    log = tmp_path / "benlsip.out"
    x_n, _, mu, pix = S.tralcnllss_native(P.x0, log_path=str(log))
    assert np.array_equal(x_h, x_n) and np.array_equal(w_h, S.fixvars_words())
    txt = log.read_text()
    assert "Outer iter 1" in txt and "AL value" in txt


def test_native_outer_loop_with_constraints_matches_host_outer_loop():
    """bnl_tralcnllss (outer loop inside the library) on the mixed-constraint family: same iterate,
    multipliers and penalty as the host-side outer loop (least_squares_multipliers, mu / omega / eta updates)."""
    P = MixedConstraintProblem(600, 24, 4)
    kw = dict(max_outer_iter=60, max_inner_iter=200)
    tr = {}
    x_h, y_h = B.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, P.A, P.b, P.xlow, P.xupp, trace=tr, **kw)
    S2 = B.Solver(0)
    S2.set_problem(P.M, P.n, P.A, P.xlow, P.xupp, p=1)
    S2.use_callbacks(P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons)
    S2.set_params(max_inner_iter=200)
    x_n, y_n, mu, pix = S2.tralcnllss_native(P.x0, max_outer_iter=60)
    assert rel(x_n, x_h) < 1e-12 and rel(y_n, y_h) < 1e-8 and mu == tr["mu"]  # y = y + mu*c amplifies rounding (mu up to 1e9)
    assert np.array_equal(S2.fixvars_words(), tr["fixvars_words"])
    S2.close()


def test_bounds_error_and_dimension_errors(S):
    S.set_problem(10, 4, None, -np.ones(4), np.ones(4))
    with pytest.raises(B.DimensionMismatch):
        S.hess_mul(np.zeros(5))
    with pytest.raises(ValueError):
        S.hess_mul(np.zeros(4))  # no Jacobian bound yet
    with pytest.raises(B.DimensionMismatch):
        S.set_problem(10, 9000)  # n > 8192 is outside the streaming kernels' range


# ---------------------------------------------------------------------------------------------------------
# Full BASELINE size (cfg3: M = 1e7, n = 1024, 81.9 GB of J): size-independent properties
# ---------------------------------------------------------------------------------------------------------
def test_full_size_properties(S):
    info = S.device_info()
    M, n = 10_000_000, 1024
    if info["free_bytes"] < 100e9:
        M = int(info["free_bytes"] * 0.6 / (8 * n))
    S.set_problem(M, n)
    S.use_builtin_model(B.MODEL_GLM, 1e-3, 0.0, 3)
    x0 = S.model_vectors()["x0"]
    S.eval_jacobian(x0)
    rng = np.random.default_rng(0)
    v, w = rng.standard_normal(n), rng.standard_normal(n)
    Hv, Hw = S.hess_mul(v), S.hess_mul(w)
    # linearity, symmetry, consistency of vthv with H*v, positive semi-definiteness
    assert rel(S.hess_mul(2.0 * v - 3.0 * w), 2.0 * Hv - 3.0 * Hw) < 1e-12
    assert abs(v @ Hw - w @ Hv) <= 1e-12 * abs(v @ Hw)
    q = S.vthv(v)
    assert q > 0 and abs(q - v @ Hv) <= 1e-12 * q
    assert np.array_equal(Hv, S.hess_mul(v))
    # x0 = 0 => z = 0, phi'(0) = 1.1, J = 1.1 A with a_ij uniform in [-1,1)/sqrt(n): diag(J'J) ~ 1.21 M / (3 n)
    e = np.zeros(n)
    e[5] = 1.0
    d = S.hess_mul(e)[5]
    assert abs(d / (1.21 * M / (3 * n)) - 1.0) < 5e-3
    # the last 1000 rows of the full-size device problem against the oracle's definition of the same rows
    tail = 1000
    P = GlmProblem(tail, n, seed=3, row0=M - tail)
    x = x0 + 0.05 * np.cos(np.arange(n))
    S.eval_jacobian(x)
    r_dev, _ = S.residuals(x)
    assert np.max(np.abs(r_dev[-tail:] - P.residuals(x))) < 1e-13
    J_tail = P.jac_res(x)
    assert np.max(np.abs(S.jv(e)[-tail:] - J_tail[:, 5])) < 1e-15
    assert rel(S.jv(v)[-tail:], J_tail @ v) < 1e-12
