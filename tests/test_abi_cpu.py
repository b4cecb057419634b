"""CPU-side checks: the C-ABI library loads and exports every symbol include/benlsip_b200.h declares (no compute
calls without a GPU), the host logic of the row sharding, and the loud failure without a device."""
import numpy as np
import pytest

import benlsip_b200 as B
from benlsip_b200.distributed import shard_rows


def test_library_exports_every_declared_symbol():
    lib = B.load_library()
    declared = B.declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/benlsip_b200.h but not exported"
    assert set(lib._bnl_signatures) == set(declared)
    assert lib.bnl_version() >= 100


def test_default_params_are_the_reference_defaults():
    """src/basic_tralcnlss.jl:177-197 and the hard-wired constants (:817, :697, :798; polyhedral_constraints.jl:207)."""
    lib = B.load_library()
    p = B.Params()
    lib.bnl_default_params(p)
    assert (p.eta1, p.eta2, p.gamma1, p.gamma2, p.kappa2, p.kappa3) == (0.25, 0.75, 0.0625, 2.0, 0.1, 0.1)
    assert p.tr_factor == 0.1 and p.atol_boundary == 1e-10
    assert p.atol_active == B.SQRT_EPS and p.atol_negcurve == B.SQRT_EPS
    assert (p.max_minor_iter, p.max_inner_iter) == (50, 500)
    o = B.OuterParams()
    lib.bnl_default_outer_params(o)
    assert (o.mu0, o.tau, o.omega0, o.eta0, o.k_crit, o.k_feas, o.beta_crit, o.beta_feas) == (10, 100, 1, 1, 1, 0.1, 1, 0.9)
    assert o.max_outer_iter == 500


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(B.NoDeviceError):
        B.Solver(0)


def test_shard_rows_partition():
    """bnl_shard_rows: whole groups of the fixed 8 x G chunk geometry (csrc/rowgeom.h) -- contiguous, disjoint, covering,
    balanced; the same chunk boundaries whatever the rank count, which is what makes the row reductions N-invariant."""
    for M, N in [(10, 1), (10, 2), (10_000_000, 8), (10_000_000, 4), (7, 8), (0, 2), (200_003, 8), (1001, 2)]:
        spans = [shard_rows(M, N, r) for r in range(N)]
        assert spans[0][0] == 0
        assert sum(m for _, m in spans) == M
        for (a0, am), (b0, _) in zip(spans, spans[1:]):
            assert a0 + am == b0
        assert max(m for _, m in spans) - min(m for _, m in spans) <= max(8 // N, 1) * 148
    # group boundaries do not move with the rank count
    b8 = [shard_rows(10_000_000, 8, r)[0] for r in range(8)]
    assert [shard_rows(10_000_000, 4, r)[0] for r in range(4)] == b8[::2]
    assert [shard_rows(10_000_000, 2, r)[0] for r in range(2)] == b8[::4]
    with pytest.raises(ValueError):
        shard_rows(10, 3, 0)  # the geometry has 8 groups: 1, 2, 4 or 8 ranks


def test_status_strings():
    lib = B.load_library()
    assert b"PosDef" in lib.bnl_status_string(-6)
    assert b"no CPU path" in lib.bnl_status_string(-9)


def test_every_entry_point_rejects_a_null_handle():
    """No entry point dereferences a NULL handle: BNL_EINVAL (-1) from every int-returning function, a message from
    bnl_last_error, nothing from bnl_destroy.  (No device needed: the guard is the first statement of every function.)"""
    import ctypes as C

    lib = B.load_library()
    checked = 0
    for name, (argtypes, restype) in sorted(lib._bnl_signatures.items()):
        if not argtypes or argtypes[0] is not C.c_void_p or name == "bnl_comm_unique_id":
            continue
        args = []
        for t in argtypes:
            if t in (C.c_void_p, C.c_char_p) or (hasattr(t, "_type_") and not isinstance(t._type_, str)):
                args.append(None)
            elif t is B.CALLBACK:
                args.append(B.CALLBACK(0))
            elif t is C.c_double:
                args.append(0.0)
            else:
                args.append(0)
        r = getattr(lib, name)(*args)
        if restype is C.c_int:
            assert r == -1, f"{name}(NULL, ...) returned {r}"
        elif restype is C.c_char_p:
            assert r
        checked += 1
    assert checked >= 45


def test_create_without_a_device_reports_enodev():
    import ctypes as C

    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = B.load_library()
    h = C.c_void_p()
    rc = lib.bnl_create(0, C.byref(h))
    assert rc == -9 and not h.value and b"no CPU path" in lib.bnl_status_string(rc)
