"""CPU statement (NumPy, against the oracle) of the algebra behind the two opt-in fast modes of the CUDA library, so the
reformulations are checked independently of any kernel:
  * incremental Cauchy search: with t = J d and u = J s_c kept up to date, phi'' = ||t||^2 and
    phi' = u.t + g.d reproduce the reference's dot(d,Hd) and dot(s_c,Hd)+dot(g,d) (src/basic_tralcnlss.jl:609-611, :633-635)
    at every breakpoint, hence the same Cauchy point and active set;
  * Gram-apply: H*v from G = J'J equals J'(Jv)."""
import numpy as np

from oracle import benlsip_oracle as O
from oracle.models import GlmProblem


def cauchy_step_incremental(x, g, H, L0, lincons, delta):
    """cauchy_step (6-arg, :574-639) for a bound-only problem with the incremental scalars."""
    n = x.shape[0]
    J = H.J
    s_c = np.zeros(n)
    O.active_bounds_reset(lincons, x, L0)
    d = O.projection(lincons, -g)
    d_u = np.minimum(lincons.xupp - x, delta)
    d_l = np.maximum(lincons.xlow - x, -delta)
    t = J @ d          # one pass
    u = np.zeros(J.shape[0])
    phi_pp = t @ t
    phi_p = u @ t + g @ d
    trace = []
    min_found = False
    while (not min_found) and lincons.nb_fix() < n:
        theta, ind = O.next_breakpoint(d, s_c, d_l, d_u, lincons.fixvars)
        trace.append((phi_p, phi_pp))
        delta_t = (-phi_p / phi_pp) if phi_pp > 0 else 0.0
        if phi_p >= 0:
            min_found = True
        elif phi_p < 0 and phi_pp > 0 and delta_t < theta:
            s_c = s_c + delta_t * d
            min_found = True
        else:
            s_c = s_c + theta * d
            u = u + theta * t                 # u = J s_c
            t = t - d[ind] * J[:, ind]        # t = J d_new : d only loses component `ind`
            O.add_active(lincons, L0, ind)
            d = O.projection(lincons, -g)
            phi_pp = t @ t
            phi_p = u @ t + g @ d
    return s_c, trace


def _literal_trace(x, g, H, L0, lincons, delta):
    """The reference's own (phi', phi'') sequence, recomputing Hd = H*d after every breakpoint."""
    n = x.shape[0]
    s_c = np.zeros(n)
    O.active_bounds_reset(lincons, x, L0)
    d = O.projection(lincons, -g)
    d_u = np.minimum(lincons.xupp - x, delta)
    d_l = np.maximum(lincons.xlow - x, -delta)
    Hd = H.mul(d)
    out = []
    while lincons.nb_fix() < n:
        phi_p, phi_pp = s_c @ Hd + g @ d, d @ Hd
        out.append((phi_p, phi_pp))
        theta, ind = O.next_breakpoint(d, s_c, d_l, d_u, lincons.fixvars)
        delta_t = (-phi_p / phi_pp) if phi_pp > 0 else 0.0
        if phi_p >= 0 or (phi_p < 0 and phi_pp > 0 and delta_t < theta):
            break
        s_c = s_c + theta * d
        O.add_active(lincons, L0, ind)
        d = O.projection(lincons, -g)
        Hd = H.mul(d)
    return out


def _state(M=1500, n=48):
    P = GlmProblem(M, n, seed=3)
    x = P.x0.copy()
    x[::7] = 1.0
    x[3::11] = -1.0
    J, r = P.jac_res(x), P.residuals(x)
    H = O.AlHessian(J, np.zeros((0, n)), 0.0)
    L0 = O._cholesky_lower(np.zeros((0, 0)))
    return P, x, J.T @ r, H, L0


def test_incremental_cauchy_scalars_equal_the_literal_ones():
    P, x, g, H, L0 = _state()
    for delta in (1e-3, 0.05, 10.0):
        c1 = O.MixedConstraints(P.A, L0, l=P.xlow, u=P.xupp)
        c2 = O.MixedConstraints(P.A, L0, l=P.xlow, u=P.xupp)
        s_lit = O.cauchy_step(x, g, H, L0, c1, delta)
        s_inc, tr_inc = cauchy_step_incremental(x, g, H, L0, c2, delta)
        np.testing.assert_allclose(s_inc, s_lit, rtol=1e-12, atol=1e-15)
        assert np.array_equal(c1.fixvars, c2.fixvars)
        c3 = O.MixedConstraints(P.A, L0, l=P.xlow, u=P.xupp)
        tr_lit = _literal_trace(x, g, H, L0, c3, delta)
        assert len(tr_lit) == len(tr_inc) and len(tr_inc) >= 1
        for (a, b), (c, d) in zip(tr_inc, tr_lit):
            assert abs(a - c) <= 1e-11 * max(abs(c), 1e-30) + 1e-9 * abs(d)  # phi' is a difference of O(phi'') terms
            assert abs(b - d) <= 1e-12 * abs(d)


def test_gram_apply_equals_matrix_free_apply():
    P, x, g, H, L0 = _state()
    G = H.J.T @ H.J
    v = np.cos(0.37 * np.arange(P.n))
    np.testing.assert_allclose(G @ v, H.mul(v), rtol=1e-12)
    assert abs(v @ G @ v - H.vthv(v)) <= 1e-12 * H.vthv(v)
