"""CPU checks of the trajectory-parity criterion (tests/parity.py) itself: which golden records count as noise-driven, and that the
well-conditioned prefix the GPU tests compare exactly is most of every solve."""
import math

from tests.parity import EPS, SQRT_EPS, first_fragile, golden


def _rec(**kw):
    r = dict(k=1, mx=100.0, delta=1.0, pix=1e-3, nb_fix=0, rho=0.9, pred=-1.0, norm_s=0.1, omega_tol=1e-6, bp_cum=0, cg_cum=0)
    r.update(kw)
    return r


def test_fragile_rules():
    assert first_fragile(dict(inner=[_rec(), _rec(k=2)])) is None
    # actual reduction of a few ulps of the AL value: rho is a ratio of rounding noise
    assert first_fragile(dict(inner=[_rec(), _rec(k=2, pred=-20 * EPS * 100.0, rho=1.2)])) == 1
    # rho comfortably away from eta1 / eta2 but |ared| large: robust
    assert first_fragile(dict(inner=[_rec(rho=0.3), _rec(k=2, rho=-5.0)])) is None
    # a trust region as small as the active-set tolerance
    assert first_fragile(dict(inner=[_rec(), _rec(k=2, delta=3.0 * SQRT_EPS)])) == 1
    # criticality measure on the subproblem tolerance
    assert first_fragile(dict(inner=[_rec(pix=1e-6 * (1 + 1e-6))])) == 0
    # NaN rho (pred == 0, trap T8) is a robust rejection
    assert first_fragile(dict(inner=[_rec(rho=float("nan"), pred=0.0)])) is None


def test_goldens_have_a_substantial_exact_prefix():
    for name in ["glm_4096_64", "glm_20000_256", "glm_6000_1024", "glm_200000_1024", "glm_500000_1024", "glm_1000000_1024", "mixed_600_24_4",
                 "mixed_4000_64_8", "mixed_20000_256_16"]:
        g = golden(name)
        F = first_fragile(g)
        n = len(g["inner"]) if F is None else F
        assert n >= 5, name
        first, last = g["inner"][0]["mx"], g["inner"][-1]["mx"]
        # the exactly-compared prefix carries (almost) all of the decrease of the AL value on the bound-constrained family
        if name.startswith("glm"):
            assert first - g["inner"][n - 1]["mx"] >= 0.999 * (first - last), name
        assert all(math.isfinite(r["mx"]) for r in g["inner"])
