"""CPU tests of the host side: (1) include/benlsip_b200.h is a valid plain-C header (C ABI: no C++ in the signatures);
(2) the host-side outer augmented-Lagrangian loop `benlsip_b200.tralcnllss` (the mirror of src/basic_tralcnlss.jl:167-298
that stays on the host) drives a subproblem solver exactly like the reference's loop -- checked with a TEST-ONLY stand-in
solver whose `solve_subproblem` is the oracle's (the product never does this: it has no CPU path)."""
import os
import subprocess
import tempfile

import numpy as np

import benlsip_b200 as B
from oracle import benlsip_oracle as O
from oracle.models import GlmProblem, MixedConstraintProblem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_compiles_as_c99():
    src = '#include "benlsip_b200.h"\nint main(void) { bnl_params p; bnl_default_params(&p); return (int)sizeof(bnl_stats) == 0; }\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", c, "-o",
                            os.path.join(d, "t.o")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_ctypes_struct_layouts_match_the_header():
    """sizeof of the POD structs as the C compiler sees them == the ctypes mirrors."""
    prog = ('#include <stdio.h>\n#include "benlsip_b200.h"\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(bnl_params), '
            'sizeof(bnl_outer_params), sizeof(bnl_stats), sizeof(bnl_inner_record)); return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        c, exe = os.path.join(d, "s.c"), os.path.join(d, "s")
        open(c, "w").write(prog)
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(t) for t in subprocess.check_output([exe], text=True).split()]
    import ctypes
    assert sizes == [ctypes.sizeof(B.Params), ctypes.sizeof(B.OuterParams), ctypes.sizeof(B.Stats), ctypes.sizeof(B.InnerRecord)]


def test_plain_c_program_links_and_calls_the_library():
    """A C99 program (no C++, no Python) links against libbenlsip_b200.so and calls it -- what a `ccall` does."""
    prog = r'''
#include <stdio.h>
#include "benlsip_b200.h"
int main(void) {
    bnl_handle h = 0;
    bnl_params p;
    bnl_default_params(&p);
    int rc = bnl_create(0, &h);
    printf("%d %d %d %s\n", bnl_version(), bnl_device_count(), rc, bnl_status_string(rc));
    if (rc == BNL_OK) bnl_destroy(h);
    return (p.max_inner_iter == 500) ? 0 : 1;
}
'''
    libdir = os.path.join(ROOT, "benlsip.jl_b200")
    with tempfile.TemporaryDirectory() as d:
        c, exe = os.path.join(d, "c.c"), os.path.join(d, "c")
        open(c, "w").write(prog)
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o", exe,
                               "-L", libdir, "-lbenlsip_b200", "-Wl,-rpath," + libdir])
        out = subprocess.check_output([exe], text=True).split(None, 3)
    version, ndev, rc = int(out[0]), int(out[1]), int(out[2])
    assert version >= 100
    if ndev == 0:
        assert rc == -9 and "no CPU path" in out[3]  # BNL_ENODEV: the library refuses to run without a B200
    else:
        assert rc in (0, -9)


class _OracleBackedSolver:
    """TEST-ONLY stand-in with the slice of the Solver interface that `tralcnllss` uses."""

    def __init__(self, P, m_lin_A):
        self.P = P
        self.n, self.p = P.n, P.nlconstraints(P.x0).shape[0]
        self.A = m_lin_A
        self.L0 = O._cholesky_lower(self.A @ self.A.T)
        self.cons = O.MixedConstraints(self.A, self.L0, l=P.xlow, u=P.xupp)
        self.kw = {}
        self.trace = {}

    def set_params(self, **kw):
        self.kw.update(kw)

    def gradient(self, x):
        return self.P.jac_res(x).T @ self.P.residuals(x)

    def nlcons(self, x):
        return self.P.nlconstraints(x), self.P.jac_nlcons(x)

    def set_fixvars(self, b):
        self.cons.fixvars[:] = b

    def solve_subproblem(self, x, y, mu, omega):
        k = self.kw
        P = self.P
        return O.solve_subproblem(x, y, mu, P.residuals, P.nlconstraints, P.jac_res, P.jac_nlcons, self.L0, self.cons,
                                  k["max_minor_iter"], k["max_inner_iter"], omega, k["eta1"], k["eta2"], k["gamma1"], k["gamma2"],
                                  k["kappa2"], k["kappa3"], trace=self.trace)

    def stats(self):
        return dict(inner_iters=self.trace.get("inner_iters", 0))

    def inner_log(self):
        return self.trace.get("inner", [])

    def fixvars_words(self):
        return self.cons.fixvars_words()

    def close(self):
        pass


def _check(P, A, **kw):
    tr_o, tr_h = {}, {}
    x_o, y_o = O.tralcnllss(P.x0, P.residuals, P.jac_res, P.nlconstraints, P.jac_nlcons, A, P.b, P.xlow, P.xupp, trace=tr_o, **kw)
    S = _OracleBackedSolver(P, A)
    x_h, y_h = B.tralcnllss(P.x0, None, None, None, None, None, None, None, None, solver=S, trace=tr_h, **kw)
    assert tr_h["outer_iters"] == tr_o["outer_iters"] and tr_h["stats"]["inner_iters"] == tr_o["inner_iters"]
    assert tr_h["mu"] == tr_o["mu"]
    np.testing.assert_allclose(x_h, x_o, rtol=0, atol=1e-13)
    np.testing.assert_allclose(y_h, y_o, rtol=1e-9, atol=1e-12)
    assert np.array_equal(tr_h["fixvars_words"], tr_o["fixvars_words"])


def test_host_outer_loop_bound_only():
    P = GlmProblem(2048, 32, seed=3)
    _check(P, P.A)


def test_host_outer_loop_with_multipliers_and_penalty_updates():
    P = MixedConstraintProblem(400, 16, 3)
    _check(P, P.A, max_outer_iter=60, max_inner_iter=200)
