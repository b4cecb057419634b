"""Property tests (hypothesis) of the oracle's building blocks: they guard the restatement the CUDA path is checked against."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import benlsip_oracle as O
from oracle.models import hash32, mix32, sym, unif


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 5), st.integers(6, 14), st.integers(0, 2 ** 31 - 1))
def test_projection_is_an_orthogonal_projector_onto_the_active_nullspace(m, n, seed):
    """projection (src/polyhedral_constraints.jl:150-170): idempotent, A v = 0, v[fix] = 0, r - v orthogonal to v."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, n))
    L = np.linalg.cholesky(A @ A.T)
    fixed = np.zeros(n, dtype=bool)
    fixed[rng.choice(n, size=rng.integers(0, n - m), replace=False)] = True
    cons = O.MixedConstraints(A, L, fixed=fixed)
    r = rng.standard_normal(n)
    v = O.projection(cons, r)
    assert np.allclose(A @ v, 0, atol=1e-9) and np.allclose(v[fixed], 0, atol=1e-10)
    assert np.allclose(O.projection(cons, v), v, atol=1e-9)
    assert abs((r - v) @ v) < 1e-8 * max(1.0, r @ r)
    # reduced-space form used by the CUDA solve path gives the same projector
    Af = A[:, ~fixed]
    vf = r[~fixed] - Af.T @ np.linalg.solve(Af @ Af.T, Af @ r[~fixed])
    assert np.allclose(v[~fixed], vf, atol=1e-8)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
def test_next_breakpoint_matches_the_sequential_scan(n, seed):
    """next_breakpoint (src/basic_tralcnlss.jl:536-562): vectorised oracle == literal strict-< scan, ties -> lowest index."""
    rng = np.random.default_rng(seed)
    d = rng.integers(-2, 3, n).astype(float)
    s = rng.integers(-1, 2, n).astype(float) * 0.5
    dl, du = -np.ones(n), np.ones(n)
    fix = rng.random(n) < 0.3
    theta, ind = np.inf, -1
    for i in range(n):
        if not fix[i]:
            t = (dl[i] - s[i]) / d[i] if d[i] < 0 else ((du[i] - s[i]) / d[i] if d[i] > 0 else np.inf)
            if t < theta:
                theta, ind = t, i
    assert O.next_breakpoint(d, s, dl, du, fix) == (theta, ind)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 200), st.integers(0, 2 ** 31 - 1))
def test_fixvars_words_are_julia_bitvector_chunks(n, seed):
    rng = np.random.default_rng(seed)
    A = np.zeros((0, n))
    cons = O.MixedConstraints(A, np.zeros((0, 0)))
    cons.fixvars[:] = rng.random(n) < 0.4
    w = cons.fixvars_words()
    assert w.shape == ((n + 63) // 64,)
    for i in range(n):
        assert bool((int(w[i >> 6]) >> (i & 63)) & 1) == bool(cons.fixvars[i])
    assert sum(bin(int(x)).count("1") for x in w) == cons.nb_fix()


def test_hash_is_the_documented_lowbias32_and_exact_in_fp64():
    # known answers of the finaliser (computed once with the C implementation in csrc/common.cuh)
    x = np.array([0, 1, 0xDEADBEEF], dtype=np.uint32)
    m = mix32(x)
    assert m[0] == 0
    y = np.uint32(1)
    y ^= y >> np.uint32(16); y = np.uint32((int(y) * 0x7FEB352D) & 0xFFFFFFFF); y ^= y >> np.uint32(15)
    y = np.uint32((int(y) * 0x846CA68B) & 0xFFFFFFFF); y ^= y >> np.uint32(16)
    assert m[1] == y
    i, j = np.arange(5, dtype=np.uint64), np.arange(7)
    h = hash32(3, i, j)
    assert h.dtype == np.uint32 and h.shape == (5, 7)
    u, s_ = unif(3, i, j), sym(3, i, j)
    assert np.all((u >= 0) & (u < 1)) and np.all((s_ >= -1) & (s_ < 1))
    assert np.array_equal(u * 2.0 ** 32, h.astype(np.float64))  # exact
    # rows beyond 2^32 wrap in the 32-bit row key by definition (M_total <= 2^32 rows)
    assert np.array_equal(hash32(3, np.array([5], dtype=np.uint64), j), hash32(3, np.array([5 + 2 ** 32], dtype=np.uint64), j))


@settings(max_examples=30, deadline=None)
@given(st.floats(1e-3, 10.0), st.floats(-2.0, 3.0))
def test_update_tr_and_nan_rho(delta, rho):
    """update_tr (src/basic_tralcnlss.jl:821-837); NaN rho leaves delta unchanged (trap T8)."""
    out = O.update_tr(delta, rho, 0.25, 0.75, 0.0625, 2.0)
    assert out == (2.0 * delta if rho > 0.75 else (0.0625 * delta if rho < 0.25 else delta))
    assert O.update_tr(delta, float("nan"), 0.25, 0.75, 0.0625, 2.0) == delta
