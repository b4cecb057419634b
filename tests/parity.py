"""Trajectory parity against the oracle's committed goldens, with the algorithm's own noise floor made explicit.

The reference's termination is generically noise-driven: once the criticality measure stalls at the rounding level of
g = J'r, the trust-region loop keeps taking steps whose actual reduction `ared = mx_next - mx` is a difference of two equal
numbers (a few ulps of mx), so `rho = ared/pred` (src/basic_tralcnlss.jl:353-354) is a ratio of rounding noise, and
`rho > eta1` / `update_tr` (:358, :821-837) flip with the summation order of ||r||^2.  No two correct FP64 implementations
(not even the reference under two BLAS thread counts) agree on those decisions.

A golden therefore carries, per inner iteration, everything needed to decide whether its decisions were numerically
meaningful (`rho`, `pred`, `mx`, `pix`, `omega_tol`).  `first_fragile(golden)` returns the first inner record whose decision
margin is below the noise threshold; the comparison is
  * EXACT up to that record (k, nb_fix, cumulative breakpoint / CG counts; mx to 1e-10, Delta to 1e-7, the pix < omega decision and the magnitude of pix);
  * if no record is fragile: exact total counts, final x to 1e-10, active-set words bit-exact;
  * otherwise the end state is compared to the noise floor (objective to 1e-12, x to 2e-8, outer count +-2), and the fragile
    record is named in the test output (-rA), so nothing is silently relaxed."""
import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EPS = float(np.finfo(np.float64).eps)
SQRT_EPS = math.sqrt(EPS)


def golden(name):
    return json.load(open(os.path.join(HERE, "golden", name + ".json")))


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def first_fragile(g, eta1=0.25, eta2=0.75, crit_tol=SQRT_EPS, tau=64.0):
    """Index of the first inner record of the golden whose accept / trust-region / termination decision is within the
    rounding noise (the margin of ared against eta1*pred and eta2*pred is below tau * eps * |mx|, i.e. |ared| is a few hundred
    ulps of mx or less; 1e-4 relative on the criticality tests), or None."""
    inner = g["inner"]
    for i, r in enumerate(inner):
        rho, pred, mx = r["rho"], r["pred"], r["mx"]
        if rho is not None and math.isfinite(rho) and math.isfinite(pred):
            ared = rho * pred
            noise = tau * EPS * abs(mx)
            if abs(ared - eta1 * pred) <= noise or abs(ared - eta2 * pred) <= noise:
                return i
        if abs(r["pix"] - r["omega_tol"]) <= 1e-4 * r["omega_tol"]:
            return i
        # a trust region as small as the active-set tolerance: active_bounds (src/polyhedral_constraints.jl:219-237) tests
        # s_i against the faces +-Delta with atol = sqrt(eps), so with Delta <= 4 sqrt(eps) WHICH variables count as active is
        # decided by the last bits of s
        if r["delta"] <= 4.0 * SQRT_EPS:
            return i
        last_of_subproblem = (i + 1 == len(inner)) or inner[i + 1]["k"] == 1
        if last_of_subproblem and abs(r["pix"] - crit_tol) <= 1e-4 * crit_tol:
            return i
    return None


def assert_trajectory_parity(name, tr_g, x_g, obj_g=None, tol=1e-10, report=print, tail_outer=2, tail_inner=16, tail_x=2e-8):
    """tr_g: trace of benlsip_b200.tralcnllss (stats, inner log, fixvars words).  Returns the fragile index (or None)."""
    g = golden(name)
    F = first_fragile(g)
    st = tr_g["stats"]
    nprefix = len(g["inner"]) if F is None else F
    assert len(tr_g["inner"]) >= nprefix
    for i in range(nprefix):
        a, b = tr_g["inner"][i], g["inner"][i]
        assert (a["k"], a["nb_fix"], a["bp_cum"], a["cg_cum"]) == (b["k"], b["nb_fix"], b["bp_cum"], b["cg_cum"]), (i, a, b)
        assert abs(a["mx"] - b["mx"]) <= tol * abs(b["mx"]), (i, a, b)
        # Delta_0 = 0.1 ||g|| and pix = ||P(-g)|| inherit the cancellation in g = J'r + C'(y + mu c) (mu up to 1e9 on the
        # mixed-constraint family): compared to 1e-7 / 1e-5, the AL value itself to `tol`
        assert abs(a["delta"] - b["delta"]) <= 1e-7 * abs(b["delta"]), (i, a, b)
        # pix = ||P(-g)|| is only ever compared with omega_tol / crit_tol, and with mu up to 1e11 the cancellation in
        # g = J'r + C'(y + mu c) leaves a small pix 1-2 significant digits: the DECISION must agree (its margin is what
        # first_fragile checks) and the magnitude (the AL value above, which does not suffer from it, is compared to `tol`)
        assert (a["pix"] < b["omega_tol"]) == (b["pix"] < b["omega_tol"]), (i, a, b)
        assert abs(a["pix"] - b["pix"]) <= 0.5 * abs(b["pix"]) + 1e-6 * b["omega_tol"] + 1e-9, (i, a, b)
    counts_g = (tr_g["outer_iters"], st["inner_iters"], st["minor_iters"], st["cg_iters"], st["breakpoints"])
    counts_o = (g["outer_iters"], g["inner_iters"], g["minor_iters"], g["cg_iters"], g["breakpoints"])
    if F is None:
        assert counts_g == counts_o
        assert rel(x_g, np.array(g["x"])) < tol
        assert [int(w) for w in tr_g["fixvars_words"]] == g["fixvars_words"]
        if obj_g is not None:
            assert abs(obj_g - g["objective"]) <= tol * g["objective"]
    else:
        r = g["inner"][F]
        why = (f"trust-region radius {r['delta']:.2e} <= 4 sqrt(eps): active-set identification is decided by the last bits of s"
               if r["delta"] <= 4.0 * SQRT_EPS else
               f"rho={r['rho']:.3g}, |ared|={abs(r['rho'] * r['pred']):.2e} vs {64 * EPS * abs(r['mx']):.2e} noise")
        report(f"[parity] {name}: golden is noise-driven from inner record {F} (k={r['k']}, {why}): exact comparison of the first "
               f"{F} records; counts cuda={counts_g} oracle={counts_o}; x rel diff {rel(x_g, np.array(g['x'])):.2e}")
        assert abs(counts_g[0] - counts_o[0]) <= tail_outer and abs(counts_g[1] - counts_o[1]) <= tail_inner
        assert rel(x_g, np.array(g["x"])) < tail_x
        if obj_g is not None:
            assert abs(obj_g - g["objective"]) <= 1e-12 * g["objective"]
    return F
