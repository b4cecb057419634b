"""
benlsip_b200 -- host-side mirror of the BEnlsip.jl interface for the inner Gauss-Newton trust-region
subproblem solve, over the C ABI of libbenlsip_b200.so (include/benlsip_b200.h).

Julia is not available in this image, so this ctypes layer is the executable stand-in for the Julia shim
(julia/BEnlsipB200.jl): same method names, argument meaning and error behaviour as the reference
(`tralcnllss` src/basic_tralcnlss.jl:167-298, `solve_subproblem` :303-378, `inner_step` :394-460,
`AlHessian` :6-10 with `*` / `vthv`, `MixedConstraints` src/polyhedral_constraints.jl:1-7 with `projection`,
`active_bounds!`, `active_bounds`, `add_active!`).  All arithmetic on the hot path runs in CUDA kernels on a
B200; there is NO CPU fallback: importing works anywhere (so symbols can be checked), creating a `Solver`
without an sm_100 GPU raises `NoDeviceError`.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbenlsip_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "benlsip_b200.h")

SQRT_EPS = math.sqrt(np.finfo(np.float64).eps)

# CG_status (src/basic_tralcnlss.jl:12); NOTHING is Julia's `nothing` (trap T3)
CG_SOLVED, CG_BOUND_HIT, CG_NEGATIVE_CURVATURE, CG_MAX_ITER, CG_NOTHING = 0, 1, 2, 3, -1
MODEL_GLM, MODEL_EXPSUM, MODEL_EXPSUM_DENSE = 1, 2, 3
HESSIAN_MATRIX_FREE, HESSIAN_GRAM = 0, 1
NLCONS_SPHERE = 1
CAUCHY_LITERAL, CAUCHY_INCREMENTAL = 0, 1


class BnlError(RuntimeError):
    pass


class NoDeviceError(BnlError):
    pass


class PosDefException(ArithmeticError):
    """LinearAlgebra.PosDefException"""


class DimensionMismatch(ValueError):
    pass


_STATUS_EXC = {-1: ValueError, -2: DimensionMismatch, -3: BnlError, -4: BnlError, -5: MemoryError, -6: PosDefException,
               -7: IndexError, -8: AssertionError, -9: NoDeviceError, -10: BnlError}


class Params(C.Structure):
    _fields_ = [("eta1", C.c_double), ("eta2", C.c_double), ("gamma1", C.c_double), ("gamma2", C.c_double),
                ("kappa2", C.c_double), ("kappa3", C.c_double), ("tr_factor", C.c_double), ("atol_active", C.c_double),
                ("atol_negcurve", C.c_double), ("atol_boundary", C.c_double), ("max_minor_iter", C.c_int32),
                ("max_inner_iter", C.c_int32)]


class OuterParams(C.Structure):
    _fields_ = [("mu0", C.c_double), ("tau", C.c_double), ("omega0", C.c_double), ("eta0", C.c_double),
                ("feas_tol", C.c_double), ("crit_tol", C.c_double), ("k_crit", C.c_double), ("k_feas", C.c_double),
                ("beta_crit", C.c_double), ("beta_feas", C.c_double), ("max_outer_iter", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("outer_iters", "inner_iters", "minor_iters", "cg_iters", "breakpoints", "hess_mul",
                                          "vthv", "jtw", "jv", "res_eval", "jac_eval", "chol_rebuilds", "allreduces")] + \
               [(k, C.c_double) for k in ("hess_mul_ms", "vthv_ms", "jtw_ms", "res_eval_ms", "jac_eval_ms", "solve_ms")] + \
               [("kernel_launches", C.c_int64), ("j_passes", C.c_int64), ("gram_count", C.c_int64), ("gram_ms", C.c_double),
                ("p2p_allreduces", C.c_int64), ("inc_breakpoints", C.c_int64), ("cauchy_loop_launches", C.c_int64),
                ("cauchy_literal_evals", C.c_int64), ("t0_reuses", C.c_int64), ("chol_downdates", C.c_int64),
                ("chol_ms", C.c_double), ("fused_jtr", C.c_int64), ("gram_breakpoints", C.c_int64), ("jt_builds", C.c_int64),
                ("point_reuses", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class InnerRecord(C.Structure):
    _fields_ = [("k", C.c_int32), ("nb_fix", C.c_int32), ("mx", C.c_double), ("norm_s", C.c_double), ("delta", C.c_double),
                ("rho", C.c_double), ("pix", C.c_double), ("pred", C.c_double), ("omega_tol", C.c_double),
                ("breakpoints_cum", C.c_int64), ("cg_cum", C.c_int64)]


CALLBACK = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)

_DP = C.POINTER(C.c_double)
_lib = None


def load_library(build_if_missing: bool = False):
    """Loads the C-ABI shared library.  Fails loudly when it is missing: there is no Python/CPU substitute."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing and os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
            from . import build as _b  # type: ignore  (compiles the CUDA sources in-tree; still no CPU substitute)
            _b.build()
        else:
            raise BnlError(f"{LIB_PATH} is missing: run `python benlsip.jl_b200/build.py` (nvcc, sm_100a). "
                           "This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    i32, i64, dbl = C.c_int32, C.c_int64, C.c_double
    sig = {
        "bnl_version": ([], C.c_int), "bnl_device_count": ([], C.c_int),
        "bnl_create": ([C.c_int, C.POINTER(H)], C.c_int), "bnl_destroy": ([H], None),
        "bnl_last_error": ([H], C.c_char_p), "bnl_status_string": ([C.c_int], C.c_char_p),
        "bnl_default_params": ([C.POINTER(Params)], None), "bnl_default_outer_params": ([C.POINTER(OuterParams)], None),
        "bnl_set_params": ([H, C.POINTER(Params)], C.c_int),
        "bnl_comm_unique_id": ([C.c_void_p], C.c_int), "bnl_comm_init": ([H, C.c_int, C.c_int, C.c_void_p], C.c_int),
        "bnl_comm_info": ([H, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)], C.c_int),
        "bnl_shard_rows": ([i64, i32, i32, C.POINTER(i64), C.POINTER(i64)], C.c_int),
        "bnl_set_problem": ([H, i64, i64, i64, i32, i32, i32, _DP, _DP, _DP], C.c_int),
        "bnl_use_builtin_model": ([H, i32, _DP, i32, C.c_uint32], C.c_int),
        "bnl_use_callbacks": ([H, CALLBACK, CALLBACK, CALLBACK, CALLBACK, C.c_void_p], C.c_int),
        "bnl_model_vectors": ([H, _DP, _DP, _DP, _DP], C.c_int),
        "bnl_use_builtin_nlcons": ([H, i32, _DP, i32], C.c_int), "bnl_model_set_truth": ([H, _DP, _DP], C.c_int),
        "bnl_nlcons": ([H, _DP, _DP, _DP], C.c_int), "bnl_gradient": ([H, _DP, _DP], C.c_int),
        "bnl_upload_jacobian": ([H, _DP, i64], C.c_int), "bnl_upload_nlcons_jacobian": ([H, _DP, i64], C.c_int),
        "bnl_set_mu": ([H, dbl], C.c_int), "bnl_eval_jacobian": ([H, _DP], C.c_int),
        "bnl_residuals": ([H, _DP, _DP, _DP], C.c_int), "bnl_hess_mul": ([H, _DP, _DP], C.c_int),
        "bnl_vthv": ([H, _DP, _DP], C.c_int), "bnl_jv": ([H, _DP, _DP], C.c_int), "bnl_jtw": ([H, _DP, _DP], C.c_int),
        "bnl_gram": ([H, _DP, _DP], C.c_int), "bnl_set_hessian_mode": ([H, i32], C.c_int), "bnl_set_cauchy_mode": ([H, i32], C.c_int), "bnl_project": ([H, _DP, _DP], C.c_int),
        "bnl_active_bounds_reset": ([H, _DP], C.c_int),
        "bnl_left_mul": ([H, _DP, _DP], C.c_int), "bnl_left_mul_tr": ([H, _DP, _DP], C.c_int),
        "bnl_active_bounds": ([H, _DP, _DP, dbl, C.POINTER(i64), C.POINTER(i32)], C.c_int),
        "bnl_add_active": ([H, C.POINTER(i64), i32], C.c_int),
        "bnl_set_fixvars": ([H, C.POINTER(C.c_uint64)], C.c_int),
        "bnl_get_fixvars": ([H, C.POINTER(C.c_uint64), C.POINTER(i32)], C.c_int),
        "bnl_get_chol": ([H, _DP, C.POINTER(i32)], C.c_int),
        "bnl_cauchy_step": ([H, _DP, _DP, dbl, _DP], C.c_int),
        "bnl_projected_cg": ([H, _DP, _DP, _DP, dbl, _DP, C.POINTER(i32), C.POINTER(i32)], C.c_int),
        "bnl_projected_cg_bounds": ([H, _DP, _DP, _DP, _DP, C.POINTER(i32), C.POINTER(i32)], C.c_int),
        "bnl_linesearch": ([H, _DP, _DP, _DP, _DP, _DP], C.c_int),
        "bnl_inner_step": ([H, _DP, _DP, dbl, _DP, _DP], C.c_int),
        "bnl_new_point": ([H, _DP, _DP, dbl, _DP, _DP, _DP], C.c_int),
        "bnl_solve_subproblem": ([H, _DP, _DP, dbl, dbl, _DP, _DP, _DP], C.c_int),
        "bnl_tralcnllss": ([H, _DP, C.POINTER(OuterParams), C.c_char_p, _DP, _DP, _DP, _DP], C.c_int),
        "bnl_get_stats": ([H, C.POINTER(Stats)], C.c_int), "bnl_reset_stats": ([H], C.c_int),
        "bnl_get_inner_log": ([H, C.POINTER(InnerRecord), i32, C.POINTER(i32)], C.c_int),
        "bnl_time_kernel": ([H, i32, i32, _DP, _DP], C.c_int),
        "bnl_device_info": ([H, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)], C.c_int),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = restype
    lib._bnl_signatures = sig
    _lib = lib
    return lib


def shard_rows(M_total: int, nranks: int, rank: int):
    """(row0, M_local) of rank `rank` of `nranks` in {1, 2, 4, 8} (`bnl_shard_rows`): whole groups of the fixed row geometry,
    so every row reduction is bit-identical for any supported GPU count.  Pure host function."""
    lib = load_library()
    r0, ml = C.c_int64(), C.c_int64()
    rc = lib.bnl_shard_rows(int(M_total), int(nranks), int(rank), C.byref(r0), C.byref(ml))
    if rc != 0:
        raise ValueError(f"bnl_shard_rows(M_total={M_total}, nranks={nranks}, rank={rank}): nranks must be 1, 2, 4 or 8")
    return int(r0.value), int(ml.value)


def declared_symbols():
    """Every function include/benlsip_b200.h declares (parsed from the header)."""
    import re
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bnl_[a-z_0-9]+)\s*\(", txt)))


def _vec(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.shape != (n,):
        raise DimensionMismatch(f"expected a vector of length {n}, got shape {a.shape}")
    return a


def _p(a):
    return a.ctypes.data_as(_DP) if a is not None else None


class Solver:
    """One solver instance bound to one GPU (one `bnl_handle`).  Holds the current (J, C, mu) = AlHessian and
    the MixedConstraints state (fixvars + factor), like the pair of mutable structs the reference passes around."""

    def __init__(self, device: int = 0):
        self.lib = load_library(build_if_missing=True)
        self.h = C.c_void_p()
        rc = self.lib.bnl_create(device, C.byref(self.h))
        if rc != 0:
            msg = self.lib.bnl_status_string(rc).decode()
            raise _STATUS_EXC.get(rc, BnlError)(f"bnl_create(device={device}) failed: {msg}")
        self.n = self.M = self.p = self.m_lin = 0
        self._cb_keep = None
        self.params = Params()
        self.lib.bnl_default_params(C.byref(self.params))

    # -- plumbing -------------------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.bnl_last_error(self.h).decode() or self.lib.bnl_status_string(rc).decode()
            raise _STATUS_EXC.get(rc, BnlError)(msg)

    def close(self):
        if self.h:
            self.lib.bnl_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup ----------------------------------------------------------------------------------------------
    def set_params(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.params, k):
                raise TypeError(f"unknown parameter {k}")
            setattr(self.params, k, v)
        self._ck(self.lib.bnl_set_params(self.h, C.byref(self.params)))

    def comm_init(self, nranks: int, rank: int, unique_id: bytes | None):
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self._ck(self.lib.bnl_comm_init(self.h, nranks, rank, buf))

    def comm_info(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self.lib.bnl_comm_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(nranks=a.value, rank=b.value, p2p_allreduce=bool(c.value))

    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.bnl_comm_unique_id(buf)
        if rc != 0:
            raise BnlError("bnl_comm_unique_id failed: " + lib.bnl_status_string(rc).decode())
        return buf.raw

    def set_problem(self, M_local, n, A=None, xlow=None, xupp=None, p=0, M_total=None, row0=0):
        """MixedConstraints(A, cholesky(A*A'); l, u) (src/polyhedral_constraints.jl:9-18, basic_tralcnlss.jl:206)."""
        A = np.zeros((0, n)) if A is None else np.asarray(A, dtype=np.float64)
        if A.ndim != 2 or A.shape[1] != n:
            raise DimensionMismatch(f"A must be m x {n}")
        m_lin = A.shape[0]
        Af = np.asfortranarray(A)
        lo = _vec(np.full(n, -np.inf) if xlow is None else xlow, n)
        up = _vec(np.full(n, np.inf) if xupp is None else xupp, n)
        self._ck(self.lib.bnl_set_problem(self.h, int(M_local), int(M_total if M_total is not None else M_local), int(row0),
                                          int(n), int(m_lin), int(p), _p(Af) if m_lin else None, _p(lo), _p(up)))
        self.n, self.M, self.p, self.m_lin = int(n), int(M_local), int(p), int(m_lin)

    def use_builtin_model(self, model_id, noise=1e-3, cond_exp=0.0, seed=3):
        prm = np.array([noise, cond_exp], dtype=np.float64)
        self._ck(self.lib.bnl_use_builtin_model(self.h, model_id, _p(prm), 2, seed))

    def use_builtin_nlcons(self, kind, rho2):
        prm = np.array([rho2], dtype=np.float64)
        self._ck(self.lib.bnl_use_builtin_nlcons(self.h, int(kind), _p(prm), 1))

    def model_set_truth(self, x_true, x0=None):
        xt = _vec(x_true, self.n)
        x0v = _vec(x0, self.n) if x0 is not None else None
        self._ck(self.lib.bnl_model_set_truth(self.h, _p(xt), _p(x0v)))

    def nlcons(self, x):
        """`nlconstraints(x)`, `jac_nlcons(x)` (built-in or callbacks) -> (c, C)."""
        c = np.zeros(self.p)
        Cm = np.zeros((self.p, self.n), order="F")
        self._ck(self.lib.bnl_nlcons(self.h, _p(_vec(x, self.n)), _p(c), _p(Cm)))
        return c, np.ascontiguousarray(Cm)

    def gradient(self, x):
        """`jac_res(x)' * residuals(x)` (src/basic_tralcnlss.jl:893)."""
        g = np.empty(self.n)
        self._ck(self.lib.bnl_gradient(self.h, _p(_vec(x, self.n)), _p(g)))
        return g

    def model_vectors(self):
        out = [np.empty(self.n) for _ in range(4)]
        self._ck(self.lib.bnl_model_vectors(self.h, *[_p(o) for o in out]))
        return dict(x0=out[0], xlow=out[1], xupp=out[2], x_true=out[3])

    def use_callbacks(self, residuals, jac_res, nlconstraints=None, jac_nlcons=None):
        """Binds the reference's four closures (src/basic_tralcnlss.jl:167-176).  Matrices cross column-major."""
        n, M, p = self.n, self.M, self.p

        def wrap(fn, shape):
            if fn is None:
                return C.cast(None, CALLBACK)

            def cb(xp, outp, _ctx):
                try:
                    x = np.ctypeslib.as_array(xp, shape=(n,)).copy()
                    val = np.asarray(fn(x), dtype=np.float64)
                    if val.shape != shape:
                        return 2
                    cnt = int(np.prod(shape))
                    if cnt:
                        np.ctypeslib.as_array(outp, shape=(cnt,))[:] = val.reshape(-1, order="F")
                    return 0
                except Exception:  # pragma: no cover
                    return 1
            return CALLBACK(cb)

        cbs = (wrap(residuals, (M,)), wrap(jac_res, (M, n)), wrap(nlconstraints, (p,)), wrap(jac_nlcons, (p, n)))
        self._cb_keep = cbs
        self._ck(self.lib.bnl_use_callbacks(self.h, *cbs, None))

    # -- AlHessian ------------------------------------------------------------------------------------------
    def upload_jacobian(self, J):
        J = np.asfortranarray(J, dtype=np.float64)
        if J.shape != (self.M, self.n):
            raise DimensionMismatch("J must be M x n")
        self._ck(self.lib.bnl_upload_jacobian(self.h, _p(J), max(self.M, 1)))

    def upload_nlcons_jacobian(self, Cm):
        Cm = np.asfortranarray(Cm, dtype=np.float64)
        if Cm.shape != (self.p, self.n):
            raise DimensionMismatch("C must be p x n")
        self._ck(self.lib.bnl_upload_nlcons_jacobian(self.h, _p(Cm), max(self.p, 1)))

    def set_mu(self, mu):
        self._ck(self.lib.bnl_set_mu(self.h, float(mu)))

    def eval_jacobian(self, x):
        self._ck(self.lib.bnl_eval_jacobian(self.h, _p(_vec(x, self.n))))

    def residuals(self, x, want_vector=True):
        r = np.empty(self.M) if want_vector else None
        ss = C.c_double()
        self._ck(self.lib.bnl_residuals(self.h, _p(_vec(x, self.n)), _p(r), C.byref(ss)))
        return r, ss.value

    def hess_mul(self, v):
        """`H*v` (src/basic_tralcnlss.jl:102-106)."""
        out = np.empty(self.n)
        self._ck(self.lib.bnl_hess_mul(self.h, _p(_vec(v, self.n)), _p(out)))
        return out

    def vthv(self, v):
        """`vthv(H,v)` (:92-96)."""
        out = C.c_double()
        self._ck(self.lib.bnl_vthv(self.h, _p(_vec(v, self.n)), C.byref(out)))
        return out.value

    def jv(self, v):
        out = np.empty(self.M)
        self._ck(self.lib.bnl_jv(self.h, _p(_vec(v, self.n)), _p(out)))
        return out

    def jtw(self, w):
        out = np.empty(self.n)
        self._ck(self.lib.bnl_jtw(self.h, _p(_vec(w, self.M)), _p(out)))
        return out

    def gram(self, want_matrix=True):
        G = np.empty((self.n, self.n), order="F") if want_matrix else None
        ms = C.c_double()
        self._ck(self.lib.bnl_gram(self.h, _p(G), C.byref(ms)))
        return G, ms.value

    def set_cauchy_mode(self, mode):
        """CAUCHY_INCREMENTAL (default: device-side guarded breakpoint loop, bit-identical Cauchy point) or CAUCHY_LITERAL
        (the reference's Hessian apply per breakpoint)."""
        self._ck(self.lib.bnl_set_cauchy_mode(self.h, int(mode)))

    def set_hessian_mode(self, mode):
        """HESSIAN_MATRIX_FREE (reference semantics, default) or HESSIAN_GRAM (G = J'J on the FP64 tensor cores)."""
        self._ck(self.lib.bnl_set_hessian_mode(self.h, int(mode)))

    # -- MixedConstraints -----------------------------------------------------------------------------------
    def projection(self, r):
        """`projection(lincons, r)` (src/polyhedral_constraints.jl:150-170)."""
        out = np.empty(self.n)
        self._ck(self.lib.bnl_project(self.h, _p(_vec(r, self.n)), _p(out)))
        return out

    def left_mul(self, x):
        """`left_mul(lincons, x)` = [A x; x[fix]] (src/polyhedral_constraints.jl:86-98)."""
        out = np.empty(self.m_lin + self.nb_fix())
        self._ck(self.lib.bnl_left_mul(self.h, _p(_vec(x, self.n)), _p(out)))
        return out

    def left_mul_tr(self, y):
        """`left_mul_tr(lincons, y)` = A~' y (src/polyhedral_constraints.jl:72-84)."""
        yv = _vec(y, self.m_lin + self.nb_fix())
        out = np.empty(self.n)
        self._ck(self.lib.bnl_left_mul_tr(self.h, _p(yv), _p(out)))
        return out

    def active_bounds_reset(self, x):
        """`active_bounds!(lincons, x, chol_aat)` (:203-215)."""
        self._ck(self.lib.bnl_active_bounds_reset(self.h, _p(_vec(x, self.n))))

    def active_bounds(self, x, s, delta):
        """`active_bounds(lincons, x, s, delta)` (:219-237) -> ascending 0-based indices."""
        idx = np.empty(self.n, dtype=np.int64)
        cnt = C.c_int32()
        self._ck(self.lib.bnl_active_bounds(self.h, _p(_vec(x, self.n)), _p(_vec(s, self.n)), float(delta),
                                            idx.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(cnt)))
        return idx[: cnt.value].copy()

    def add_active(self, ind):
        """`add_active!` (:240-261), Int or Vector{Int} (0-based)."""
        idx = np.atleast_1d(np.asarray(ind, dtype=np.int64))
        self._ck(self.lib.bnl_add_active(self.h, idx.ctypes.data_as(C.POINTER(C.c_int64)), int(idx.size)))

    def set_fixvars(self, fixed_bool):
        b = np.asarray(fixed_bool, dtype=bool)
        nw = (self.n + 63) // 64
        bits = np.zeros(nw * 64, dtype=np.uint8)
        bits[: self.n] = b
        words = np.packbits(bits.reshape(nw, 64), axis=1, bitorder="little").view(np.uint64).reshape(nw).copy()
        self._ck(self.lib.bnl_set_fixvars(self.h, words.ctypes.data_as(C.POINTER(C.c_uint64))))

    def fixvars_words(self):
        """`lincons.fixvars.chunks` (UInt64 words, LSB first)."""
        nw = (self.n + 63) // 64
        words = np.zeros(nw, dtype=np.uint64)
        cnt = C.c_int32()
        self._ck(self.lib.bnl_get_fixvars(self.h, words.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(cnt)))
        return words

    def fixvars(self):
        w = self.fixvars_words()
        return np.unpackbits(w.view(np.uint8), bitorder="little")[: self.n].astype(bool)

    def nb_fix(self):
        return int(self.fixvars().sum())

    def chol_L(self):
        dim = C.c_int32()
        self._ck(self.lib.bnl_get_chol(self.h, None, C.byref(dim)))
        L = np.zeros((dim.value, dim.value), order="F")
        self._ck(self.lib.bnl_get_chol(self.h, _p(L), C.byref(dim)))
        return np.ascontiguousarray(L)

    # -- step computation -----------------------------------------------------------------------------------
    def cauchy_step(self, x, g, delta):
        out = np.empty(self.n)
        self._ck(self.lib.bnl_cauchy_step(self.h, _p(_vec(x, self.n)), _p(_vec(g, self.n)), float(delta), _p(out)))
        return out

    def projected_cg(self, x, s, g_minor, delta):
        w = np.empty(self.n)
        st, it = C.c_int32(), C.c_int32()
        self._ck(self.lib.bnl_projected_cg(self.h, _p(_vec(x, self.n)), _p(_vec(s, self.n)), _p(_vec(g_minor, self.n)),
                                           float(delta), _p(w), C.byref(st), C.byref(it)))
        return w, (None if st.value == CG_NOTHING else st.value), it.value

    def projected_cg_bounds(self, g_minor, w_l, w_u):
        """`projected_cg(g_minor, H, w_l, w_u, lincons, kappa2)` (:690-764) with explicit step bounds and the current fixvars."""
        w = np.empty(self.n)
        st, it = C.c_int32(), C.c_int32()
        self._ck(self.lib.bnl_projected_cg_bounds(self.h, _p(_vec(g_minor, self.n)), _p(_vec(w_l, self.n)), _p(_vec(w_u, self.n)),
                                                  _p(w), C.byref(st), C.byref(it)))
        return w, (None if st.value == CG_NOTHING else st.value), it.value

    def linesearch(self, g_model, w, w_l, w_u):
        """`linesearch(g_model, H, w, w_l, w_u, lincons.fixvars)` (:766-791) -> alpha."""
        out = C.c_double()
        self._ck(self.lib.bnl_linesearch(self.h, _p(_vec(g_model, self.n)), _p(_vec(w, self.n)), _p(_vec(w_l, self.n)),
                                         _p(_vec(w_u, self.n)), C.byref(out)))
        return out.value

    def inner_step(self, x, g, delta):
        """`inner_step(...)` (:394-460) -> (s, model_reduction); mutates the active set."""
        s = np.empty(self.n)
        pred = C.c_double()
        self._ck(self.lib.bnl_inner_step(self.h, _p(_vec(x, self.n)), _p(_vec(g, self.n)), float(delta), _p(s), C.byref(pred)))
        return s, pred.value

    def new_point(self, x, y, mu):
        """`new_point` (:32-49) -> (mx, g, cx); J, C, mu stay in the handle as H."""
        mx = C.c_double()
        g = np.empty(self.n)
        cx = np.empty(self.p)
        yv = _vec(y if y is not None else np.zeros(self.p), self.p)
        self._ck(self.lib.bnl_new_point(self.h, _p(_vec(x, self.n)), _p(yv), float(mu), C.byref(mx), _p(g), _p(cx)))
        return mx.value, g, cx

    def solve_subproblem(self, x0, y, mu, omega_tol):
        """`solve_subproblem` (:303-378) -> (x, cx, pix).  TR / CG parameters come from `set_params`."""
        x = np.empty(self.n)
        cx = np.empty(self.p)
        pix = C.c_double()
        yv = _vec(y if y is not None else np.zeros(self.p), self.p)
        self._ck(self.lib.bnl_solve_subproblem(self.h, _p(_vec(x0, self.n)), _p(yv), float(mu), float(omega_tol), _p(x),
                                               _p(cx), C.byref(pix)))
        return x, cx, pix.value

    def tralcnllss_native(self, x0, log_path=None, **outer_kw):
        """Outer loop inside the library (`bnl_tralcnllss`)."""
        op = OuterParams()
        self.lib.bnl_default_outer_params(C.byref(op))
        for k, v in outer_kw.items():
            setattr(op, k, v)
        x, y = np.empty(self.n), np.empty(self.p)
        mu, pix = C.c_double(), C.c_double()
        self._ck(self.lib.bnl_tralcnllss(self.h, _p(_vec(x0, self.n)), C.byref(op),
                                         log_path.encode() if log_path else None, _p(x), _p(y), C.byref(mu), C.byref(pix)))
        return x, y, mu.value, pix.value

    # -- introspection --------------------------------------------------------------------------------------
    def stats(self):
        s = Stats()
        self._ck(self.lib.bnl_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._ck(self.lib.bnl_reset_stats(self.h))

    def inner_log(self):
        cnt = C.c_int32()
        self._ck(self.lib.bnl_get_inner_log(self.h, None, 0, C.byref(cnt)))
        arr = (InnerRecord * max(cnt.value, 1))()
        self._ck(self.lib.bnl_get_inner_log(self.h, arr, cnt.value, C.byref(cnt)))
        return [dict(k=a.k, nb_fix=a.nb_fix, mx=a.mx, norm_s=a.norm_s, delta=a.delta, rho=a.rho, pix=a.pix, pred=a.pred,
                     omega_tol=a.omega_tol, bp_cum=a.breakpoints_cum, cg_cum=a.cg_cum)
                for a in arr[: cnt.value]]

    def time_kernel(self, kind, reps=10):
        ms, nbytes = C.c_double(), C.c_double()
        self._ck(self.lib.bnl_time_kernel(self.h, kind, reps, C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value

    def device_info(self):
        sm, maj, mi = C.c_int32(), C.c_int32(), C.c_int32()
        fr, tot = C.c_int64(), C.c_int64()
        self._ck(self.lib.bnl_device_info(self.h, C.byref(sm), C.byref(maj), C.byref(mi), C.byref(fr), C.byref(tot)))
        return dict(sm_count=sm.value, cc=(maj.value, mi.value), free_bytes=fr.value, total_bytes=tot.value)


# ==========================================================================================================
# tralcnllss: the reference's only export (src/basic_tralcnlss.jl:167-298).  The outer augmented-Lagrangian
# loop stays on the host (here Python standing in for Julia) and calls the library once per outer iteration.
# ==========================================================================================================
def initial_tolerances(mu, omega0, eta0, k_crit, k_feas):
    """src/basic_tralcnlss.jl:153-163."""
    return omega0 / (mu ** k_crit), eta0 / (mu ** k_feas)


def first_order_multipliers(y, cx, mu):
    """src/basic_tralcnlss.jl:905-911."""
    return y + mu * cx


def tralcnllss(x0, residuals, jac_res, nlconstraints, jac_nlcons, A, b, x_l, x_u, *,
               mu0=10.0, tau=100.0, omega0=1.0, eta0=1.0, feas_tol=SQRT_EPS, crit_tol=SQRT_EPS,
               k_crit=1.0, k_feas=0.1, beta_crit=1.0, beta_feas=0.9, eta1=0.25, eta2=0.75,
               gamma1=0.0625, gamma2=2.0, gamma_c=10.0, kappa1=1e-2, kappa2=0.1, kappa3=0.1,
               max_outer_iter=500, max_inner_iter=500, max_minor_iter=50,
               solver: Solver | None = None, device=0, trace: dict | None = None):
    """Same signature, keywords and defaults as the reference.  Returns (x, y).

    `residuals` etc. are host callables (uploaded through pinned memory on every evaluation), or all None when
    `solver` already has a built-in device model bound (`Solver.use_builtin_model`)."""
    assert (0 < eta1 <= eta2 < 1) and (0 < gamma1 < 1 < gamma2), "Invalid trust region updates paramaters"
    x = np.array(x0, dtype=np.float64, copy=True)
    n = x.shape[0]
    builtin = residuals is None
    own = solver is None
    if own:
        if builtin:
            raise ValueError("a built-in model needs a prepared Solver")
        A = np.asarray(A, dtype=np.float64).reshape(-1, n)
        r0 = np.asarray(residuals(x), dtype=np.float64)
        c0 = np.asarray(nlconstraints(x), dtype=np.float64)
        solver = Solver(device)
        solver.set_problem(r0.shape[0], n, A, x_l, x_u, p=c0.shape[0])  # chol_aat = cholesky(A*A') :206
        solver.use_callbacks(residuals, jac_res, nlconstraints, jac_nlcons)
    S = solver
    S.set_params(eta1=eta1, eta2=eta2, gamma1=gamma1, gamma2=gamma2, kappa2=kappa2, kappa3=kappa3,
                 max_minor_iter=max_minor_iter, max_inner_iter=max_inner_iter)
    p = S.p
    mu = float(mu0)
    omega, eta = initial_tolerances(mu0, omega0, eta0, k_crit, k_feas)  # :229
    # least_squares_multipliers :887-903
    if p > 0:
        g = S.gradient(x)  # jac_res(x)' * residuals(x) on the device
        Cm = S.nlcons(x)[1] if builtin else np.asarray(jac_nlcons(x), dtype=np.float64)
        L = np.linalg.cholesky(Cm @ Cm.T)
        y = np.linalg.solve(L.T, np.linalg.solve(L, -(Cm @ g)))
    else:
        y = np.zeros(0)
    S.set_fixvars(np.zeros(n, dtype=bool))  # MixedConstraints(A, chol_aat; l, u) :231
    cx = np.zeros(p)
    first_order_critical = False
    outer_iter = 1
    pix = math.inf
    while (not first_order_critical) and outer_iter <= max_outer_iter:  # :246
        x_next, cx_next, pix = S.solve_subproblem(x, y, mu, omega)  # :249-268  <-- the C-ABI boundary
        feas_measure = float(np.linalg.norm(cx_next))
        if feas_measure <= eta:  # :273
            x[:] = x_next
            cx = cx_next.copy()
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
        if trace is not None:
            trace.setdefault("outer", []).append(dict(outer_iter=outer_iter, feas=feas_measure, mu=mu, pix=pix, omega=omega))
    if trace is not None:
        trace["outer_iters"] = outer_iter - 1
        trace["stats"] = S.stats()
        trace["inner"] = S.inner_log()
        trace["fixvars_words"] = S.fixvars_words()
        trace["mu"] = mu
        trace["pix"] = pix
    if own:
        S.close()
    return x, y
