"""The three descriptions of the boundary must agree: include/benlsip_b200.h (the contract), the ctypes mirror
(benlsip.jl_b200/__init__.py, executed by every test) and the Julia shim (julia/BEnlsipB200.jl, which cannot be executed
here: no Julia in the image).  The header is parsed, every prototype is compared with the ctypes argtypes/restype and with
every `ccall` of the shim (argument count and the C type each Julia type lowers to), and the by-value structs are compared
field by field (order, type, offsets through ctypes).  CPU only: nothing is computed."""
import ctypes as C
import os
import re

import benlsip_b200 as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "benlsip_b200.h")
JULIA = os.path.join(ROOT, "julia", "BEnlsipB200.jl")


def _header_text():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    return re.sub(r"//[^\n]*", " ", txt)


def _canon(ctype: str) -> str:
    """'const double* x' -> 'double*', 'bnl_handle h' -> 'handle', 'int32_t n' -> 'i32' ..."""
    t = ctype.strip()
    t = re.sub(r"\bconst\b", "", t)
    stars = t.count("*")
    t = t.replace("*", " ")
    words = t.split()
    if len(words) > 1 and words[-1] not in ("int", "double", "void", "char"):
        words = words[:-1]  # drop the parameter name
    base = " ".join(words)
    base = {"int": "i32", "int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "double": "f64",
            "void": "void", "char": "char", "bnl_handle": "handle", "bnl_callback": "fnptr"}.get(base, base)
    return base + "*" * stars


def header_prototypes():
    protos = {}
    for ret, name, args in re.findall(r"\b(int|void|const\s+char\s*\*)\s+(bnl_[a-z_0-9]+)\s*\(([^()]*)\)\s*;", _header_text()):
        args = args.strip()
        argl = [] if args in ("", "void") else [_canon(a) for a in args.split(",")]
        protos[name] = (argl, _canon(ret))
    return protos


def header_struct(name):
    body = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), _header_text(), flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ty, names = decl.split(None, 1)
        for nm in names.split(","):
            fields.append((nm.strip(), _canon(ty)))
    return fields


_CT = {C.c_int: "i32", C.c_int32: "i32", C.c_uint32: "u32", C.c_int64: "i64", C.c_uint64: "u64", C.c_double: "f64",
       C.c_char_p: "char*", None: "void"}


def _ctypes_canon(t, proto_hint=None):
    if t in _CT:
        return _CT[t]
    if t is C.c_void_p:
        return proto_hint if proto_hint in ("handle", "void*") else "void*"
    if t is B.CALLBACK:
        return "fnptr"
    if hasattr(t, "_type_") and not isinstance(t._type_, str):  # POINTER(x)
        inner = t._type_
        if inner is C.c_void_p:
            return "handle*"
        if issubclass(inner, C.Structure):
            return {"Params": "bnl_params", "OuterParams": "bnl_outer_params", "Stats": "bnl_stats",
                    "InnerRecord": "bnl_inner_record"}[inner.__name__] + "*"
        return _CT[inner] + "*"
    raise AssertionError(f"unmapped ctypes type {t}")


def test_header_parses_completely():
    protos = header_prototypes()
    assert set(protos) == set(B.declared_symbols()), set(B.declared_symbols()) ^ set(protos)


def test_ctypes_signatures_equal_the_header():
    lib = B.load_library()
    protos = header_prototypes()
    for name, (argtypes, restype) in lib._bnl_signatures.items():
        hargs, hret = protos[name]
        assert len(argtypes) == len(hargs), f"{name}: {len(argtypes)} ctypes arguments, header has {len(hargs)}"
        got = [_ctypes_canon(t, h) for t, h in zip(argtypes, hargs)]
        assert got == hargs, f"{name}: ctypes {got} != header {hargs}"
        assert _ctypes_canon(restype) == hret, f"{name}: restype"


def test_ctypes_structs_equal_the_header():
    for cls, cname in [(B.Params, "bnl_params"), (B.OuterParams, "bnl_outer_params"), (B.Stats, "bnl_stats"),
                       (B.InnerRecord, "bnl_inner_record")]:
        hf = header_struct(cname)
        cf = [(n, _CT[t]) for n, t in cls._fields_]
        assert cf == hf, f"{cname}: {cf} != {hf}"
        # natural alignment, no surprises: every 8-byte field sits on an 8-byte offset, the size is a multiple of 8
        for n, t in cls._fields_:
            if C.sizeof(t) == 8:
                assert getattr(cls, n).offset % 8 == 0
        assert C.sizeof(cls) % 8 == 0


_JL = {"Cint": "i32", "Int32": "i32", "Int64": "i64", "Cdouble": "f64", "Cstring": "char*", "Cvoid": "void",
       "Ptr{Cdouble}": "f64*", "Ref{Cdouble}": "f64*", "Ptr{UInt64}": "u64*", "Ptr{Int64}": "i64*", "Ref{Int32}": "i32*",
       "Ref{BnlParams}": "bnl_params*", "Ptr{InnerRecord}": "bnl_inner_record*", "Ptr{Ptr{Cvoid}}": "handle*"}


def julia_ccalls():
    src = open(JULIA).read()
    src = re.sub(r"#[^\n]*", "", src)
    calls = []
    for m in re.finditer(r"ccall\(\(:(bnl_[a-z_0-9]+),\s*LIB\),\s*(\w+),\s*\(([^()]*)\)\s*,?", src):
        name, ret, argt = m.group(1), m.group(2), m.group(3)
        args = [a.strip() for a in argt.split(",") if a.strip()]
        # count the values passed after the type tuple (balanced scan to the closing parenthesis of ccall)
        i, depth, nvals, cur = m.end(), 1, 0, ""
        while depth > 0:
            ch = src[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    break
            if ch == "," and depth == 1:
                nvals += bool(cur.strip())
                cur = ""
            else:
                cur += ch
            i += 1
        nvals += bool(cur.strip())
        calls.append((name, ret, args, nvals))
    return calls


def test_julia_ccalls_match_the_header():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 25
    for name, ret, args, nvals in calls:
        assert name in protos, f"{name} is not declared in the header"
        hargs, hret = protos[name]
        assert len(args) == len(hargs) == nvals, f"{name}: {len(args)} Julia types, {nvals} values, header has {len(hargs)}"
        assert _JL[ret] == hret, f"{name}: return type {ret}"
        for jt, ht in zip(args, hargs):
            if jt == "Ptr{Cvoid}":
                assert ht in ("handle", "void*", "fnptr"), f"{name}: Ptr{{Cvoid}} passed where the header has {ht}"
            else:
                assert _JL[jt] == ht, f"{name}: Julia {jt} where the header has {ht}"
    # every reference method the shim claims to forward is really bound
    used = {c[0] for c in calls}
    for must in ("bnl_solve_subproblem", "bnl_inner_step", "bnl_hess_mul", "bnl_vthv", "bnl_project", "bnl_active_bounds_reset",
                 "bnl_active_bounds", "bnl_add_active", "bnl_set_fixvars", "bnl_get_fixvars", "bnl_get_chol",
                 "bnl_use_callbacks", "bnl_upload_jacobian", "bnl_upload_nlcons_jacobian", "bnl_set_mu", "bnl_get_inner_log"):
        assert must in used, must


def test_julia_structs_equal_the_header():
    src = open(JULIA).read()
    jl_ty = {"Cdouble": "f64", "Int32": "i32", "Int64": "i64"}
    for jname, cname in [("BnlParams", "bnl_params"), ("InnerRecord", "bnl_inner_record")]:
        body = re.search(r"struct %s\n(.*?)\nend" % jname, src, flags=re.S).group(1)
        fields = [(f.split("::")[0].strip(), jl_ty[f.split("::")[1].strip()])
                  for line in body.splitlines() for f in line.split(";") if "::" in f]
        assert fields == header_struct(cname), f"{jname}: {fields}"


def test_julia_overrides_use_the_reference_arities():
    """enable!() installs methods with the reference's own positional arities: solve_subproblem 18 (src/basic_tralcnlss.jl:303-322),
    inner_step 9 (:394-404), add_active! 3, active_bounds 4, active_bounds! 3, projection! 3."""
    src = open(JULIA).read()
    block = src[src.index("@eval BEnlsip begin"):]

    def arity(fname):
        m = re.search(r"\n\s+%s\(" % re.escape(fname), block)
        i, depth, n, kw = m.end(), 1, 1, False
        while depth:
            ch = block[i]
            depth += ch in "([{"
            depth -= ch in ")]}"
            if depth == 1 and ch == ";":
                kw = True
            if depth == 1 and ch == "," and not kw:
                n += 1
            i += 1
        return n

    assert arity("solve_subproblem") == 18
    assert arity("inner_step") == 9
    assert arity("add_active!") == 3
    assert arity("active_bounds") == 4
    assert arity("active_bounds!") == 3
    assert arity("projection!") == 3
    assert arity("projection") == 2
