# BEnlsipB200.jl -- Julia shim: BEnlsip.jl's OWN methods, on its OWN structs, forwarded to libbenlsip_b200.so
# (include/benlsip_b200.h) through `ccall`.  After
#
#     using BEnlsip; include("julia/BEnlsipB200.jl"); BEnlsipB200.enable!()
#
# the package's solver entry point `tralcnllss` (src/basic_tralcnlss.jl:167-298) runs unchanged -- the augmented-Lagrangian
# outer loop stays Julia -- while the hot path below it runs on the B200:
#
#   solve_subproblem (18 positional arguments, :303-322)          -> bnl_solve_subproblem
#   inner_step(x,g,H,chol_aat,lincons,delta,nb_minor_step,kappa2,kappa3) (:394-404) -> bnl_inner_step
#   Base.:*(H::AlHessian, v) (:102-106), vthv(H, v) (:92-96)      -> bnl_hess_mul, bnl_vthv
#   projection(lincons, r) / projection!(lincons, r, v) (src/polyhedral_constraints.jl:150-170) -> bnl_project
#   active_bounds!(lincons, x, chol_aat) (:203-215)               -> bnl_active_bounds_reset
#   active_bounds(lincons, x, s, delta) (:219-237)                -> bnl_active_bounds
#   add_active!(lincons, chol_aat, ind | indx) (:240-261)         -> bnl_add_active
#
# Every method that mutates `lincons` in the reference writes `lincons.fixvars` (BitVector chunks cross the ABI unchanged)
# and `lincons.chol` (the (m+q)^2 factor, `bnl_get_chol`) back, so Julia code that inspects the structs afterwards sees what the
# reference would have left there.
#
# NOT EXECUTED in this repository's CI: there is no Julia in the build image or on the GPU box.  The tested host side is its
# transliteration benlsip.jl_b200/__init__.py (same calls, same order); INTEGRATION.md shows the correspondence.
module BEnlsipB200

using LinearAlgebra
import BEnlsip
import BEnlsip: AlHessian, MixedConstraints

const LIB = get(ENV, "BENLSIP_B200_LIB", joinpath(@__DIR__, "..", "benlsip.jl_b200", "libbenlsip_b200.so"))

struct BnlParams
    eta1::Cdouble; eta2::Cdouble; gamma1::Cdouble; gamma2::Cdouble
    kappa2::Cdouble; kappa3::Cdouble
    tr_factor::Cdouble; atol_active::Cdouble; atol_negcurve::Cdouble; atol_boundary::Cdouble
    max_minor_iter::Int32; max_inner_iter::Int32
end

# status code -> the exception the reference would have thrown (SURVEY.md section 5)
function check(h::Ptr{Cvoid}, rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:bnl_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    rc == -6 && throw(PosDefException(0))
    rc == -7 && throw(BoundsError())
    rc == -8 && throw(AssertionError(msg))
    rc == -2 && throw(DimensionMismatch(msg))
    error("libbenlsip_b200: $msg (status $rc)")
end

# ---- one device handle per MixedConstraints object (the reference threads `lincons` through every call) -----------------------
mutable struct Solver
    h::Ptr{Cvoid}
    n::Int; M::Int; p::Int; m::Int
    cbs::Any                     # keeps the @cfunction closures alive
    closures::Any                # (residuals, jac_res, nlconstraints, jac_nlcons) currently bound
    J_ref::Any; C_ref::Any       # the AlHessian matrices currently on the device; holding them keeps `===` meaningful (a freed
                                 # matrix's address can be reused by the next jac_res(x)); in-place edits of H.J need rebind!(H)
    mu::Float64
    function Solver(device::Integer=0)
        hp = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:bnl_create, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), device, hp)
        rc == 0 || error("bnl_create: " * unsafe_string(ccall((:bnl_status_string, LIB), Cstring, (Cint,), rc)))
        s = new(hp[], 0, 0, 0, 0, nothing, nothing, nothing, nothing, NaN)
        finalizer(s -> ccall((:bnl_destroy, LIB), Cvoid, (Ptr{Cvoid},), s.h), s)
        return s
    end
end

const DEVICE = Ref(0)
# lincons (or an AlHessian used on its own) -> its handle.  Weak keys: both structs are mutable (identity-hashed), and a handle --
# device buffers included -- is released by its finalizer once the struct it served is garbage
const SOLVERS = WeakKeyDict{Any,Solver}()

# MixedConstraints(A, cholesky(A*A'); l, u)   src/polyhedral_constraints.jl:9-18, src/basic_tralcnlss.jl:206
function set_problem!(s::Solver, M::Integer, lincons::MixedConstraints{Float64}, p::Integer)
    m, n = size(lincons.lineq)
    A, x_l, x_u = lincons.lineq, lincons.xlow, lincons.xupp
    GC.@preserve A x_l x_u check(s.h, ccall((:bnl_set_problem, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Int64, Int32, Int32, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        s.h, M, M, 0, n, m, p, m == 0 ? C_NULL : pointer(A), pointer(x_l), pointer(x_u)))
    s.n, s.M, s.p, s.m = n, M, p, m
    s.J_ref = s.C_ref = nothing
end

# the handle of `lincons`, created (and dimensioned for M residuals, p nonlinear constraints) on first use
function solver_for(lincons::MixedConstraints{Float64}, M::Integer, p::Integer)
    s = get!(() -> Solver(DEVICE[]), SOLVERS, lincons)
    if s.M != M || s.p != p || s.n != size(lincons.lineq, 2) || s.m != size(lincons.lineq, 1)
        set_problem!(s, M, lincons, p)
    end
    push_fixvars!(s, lincons)
    return s
end

# lincons.fixvars -> device (update_chol! happens there), and device -> lincons.fixvars / lincons.chol
push_fixvars!(s::Solver, lincons) =
    check(s.h, ccall((:bnl_set_fixvars, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}), s.h, lincons.fixvars.chunks))
function pull_lincons!(s::Solver, lincons::MixedConstraints{Float64})
    cnt = Ref{Int32}(0)
    check(s.h, ccall((:bnl_get_fixvars, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ref{Int32}), s.h, lincons.fixvars.chunks, cnt))
    dim = Ref{Int32}(0)
    check(s.h, ccall((:bnl_get_chol, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ref{Int32}), s.h, C_NULL, dim))
    L = Matrix{Float64}(undef, dim[], dim[])
    check(s.h, ccall((:bnl_get_chol, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ref{Int32}), s.h, L, dim))
    lincons.chol = Cholesky(L, 'L', 0)      # lincons.chol.L is what the reference's tests compare (test/structures.jl:33)
    return lincons
end

# the four closures of tralcnllss (src/basic_tralcnlss.jl:167-176) as C callbacks; matrices cross column-major
function use_callbacks!(s::Solver, residuals, jac_res, nlconstraints, jac_nlcons)
    s.closures === (residuals, jac_res, nlconstraints, jac_nlcons) && return
    n = s.n
    mk(f, len) = begin
        function cb(xp::Ptr{Cdouble}, outp::Ptr{Cdouble}, ::Ptr{Cvoid})::Cint
            x = copy(unsafe_wrap(Array, xp, n))
            val = f(x)
            length(val) == len || return Cint(2)
            len > 0 && copyto!(unsafe_wrap(Array, outp, len), vec(val))
            return Cint(0)
        end
        @cfunction($cb, Cint, (Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}))
    end
    cbs = (mk(residuals, s.M), mk(jac_res, s.M * n), mk(nlconstraints, s.p), mk(jac_nlcons, s.p * n))
    s.cbs = cbs
    s.closures = (residuals, jac_res, nlconstraints, jac_nlcons)
    check(s.h, ccall((:bnl_use_callbacks, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                     s.h, cbs[1], cbs[2], cbs[3], cbs[4], C_NULL))
    s.J_ref = s.C_ref = nothing
end

function set_params!(s::Solver; eta1=0.25, eta2=0.75, gamma1=0.0625, gamma2=2.0, kappa2=0.1, kappa3=0.1,
                     max_minor_iter=50, max_inner_iter=500)
    p = Ref(BnlParams(eta1, eta2, gamma1, gamma2, kappa2, kappa3, 0.1, sqrt(eps()), sqrt(eps()), 1e-10,
                      max_minor_iter, max_inner_iter))
    check(s.h, ccall((:bnl_set_params, LIB), Cint, (Ptr{Cvoid}, Ref{BnlParams}), s.h, p))
end

# modes (INTEGRATION.md 2b)
const HESSIAN_MATRIX_FREE, HESSIAN_GRAM = Int32(0), Int32(1)
const CAUCHY_LITERAL, CAUCHY_INCREMENTAL = Int32(0), Int32(1)
set_hessian_mode!(s::Solver, mode::Int32) = check(s.h, ccall((:bnl_set_hessian_mode, LIB), Cint, (Ptr{Cvoid}, Int32), s.h, mode))
set_cauchy_mode!(s::Solver, mode::Int32) = check(s.h, ccall((:bnl_set_cauchy_mode, LIB), Cint, (Ptr{Cvoid}, Int32), s.h, mode))

# ---- AlHessian (src/basic_tralcnlss.jl:6-10): (J, C, mu) uploaded through pinned staging when the struct's matrices change ----
function bind_hessian!(s::Solver, H::AlHessian{Float64})
    if H.J !== s.J_ref
        check(s.h, ccall((:bnl_upload_jacobian, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64), s.h, H.J, size(H.J, 1)))
        s.J_ref = H.J
    end
    if s.p > 0 && H.C !== s.C_ref
        check(s.h, ccall((:bnl_upload_nlcons_jacobian, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64), s.h, H.C, size(H.C, 1)))
        s.C_ref = H.C
    end
    if H.mu != s.mu
        check(s.h, ccall((:bnl_set_mu, LIB), Cint, (Ptr{Cvoid}, Cdouble), s.h, H.mu))
        s.mu = H.mu
    end
    return s
end
# after H.J or H.C was modified in place: forget the cached matrices so the next call uploads them again
rebind!(H::AlHessian{Float64}) = (haskey(SOLVERS, H) && (SOLVERS[H].J_ref = SOLVERS[H].C_ref = nothing); H)
# an AlHessian used on its own (test/structures.jl:1-16): a handle without constraints, keyed by the struct
function solver_for(H::AlHessian{Float64})
    M, n = size(H.J); p = size(H.C, 1)
    s = get!(() -> Solver(DEVICE[]), SOLVERS, H)
    if s.M != M || s.n != n || s.p != p
        lo, up = fill(-Inf, n), fill(Inf, n)
        check(s.h, ccall((:bnl_set_problem, LIB), Cint,
            (Ptr{Cvoid}, Int64, Int64, Int64, Int32, Int32, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
            s.h, M, M, 0, n, 0, p, C_NULL, lo, up))
        s.n, s.M, s.p, s.m = n, M, p, 0
        s.J_ref = s.C_ref = nothing
    end
    return bind_hessian!(s, H)
end

hess_mul(s::Solver, v::Vector{Float64}) = (out = similar(v);
    check(s.h, ccall((:bnl_hess_mul, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.h, v, out)); out)
vthv(s::Solver, v::Vector{Float64}) = (out = Ref{Cdouble}(0);
    check(s.h, ccall((:bnl_vthv, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ref{Cdouble}), s.h, v, out)); out[])

# ---- the reference's methods, same signatures, on its own structs -----------------------------------------------------------
# Base.:*(H::AlHessian, v) :102-106 ; vthv(H, v) :92-96
b200_mul(H::AlHessian{Float64}, v::Vector{Float64}) = hess_mul(solver_for(H), v)
b200_vthv(H::AlHessian{Float64}, v::Vector{Float64}) = vthv(solver_for(H), v)

# projection(lincons, r) / projection!(lincons, r, v)   src/polyhedral_constraints.jl:150-170
function b200_projection!(lincons::MixedConstraints{Float64}, r::Vector{Float64}, v::Vector{Float64})
    s = lincons_solver(lincons)
    check(s.h, ccall((:bnl_project, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.h, r, v))
    return
end
b200_projection(lincons::MixedConstraints{Float64}, r::Vector{Float64}) = (v = similar(r); b200_projection!(lincons, r, v); v)

# active_bounds!(lincons, x, chol_aat; atol) :203-215 -- overwrites fixvars from x, rebuilds the factor
function b200_active_bounds!(lincons::MixedConstraints{Float64}, x::Vector{Float64}, chol_aat::Cholesky; atol::Float64=sqrt(eps()))
    s = lincons_solver(lincons)
    check(s.h, ccall((:bnl_active_bounds_reset, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.h, x))
    pull_lincons!(s, lincons)
    return
end
# active_bounds(lincons, x, s, delta; atol) :219-237 -> ascending 1-based indices
function b200_active_bounds(lincons::MixedConstraints{Float64}, x::Vector{Float64}, st::Vector{Float64}, delta::Float64;
                            atol::Float64=sqrt(eps()))
    s = lincons_solver(lincons)
    idx = Vector{Int64}(undef, length(x)); cnt = Ref{Int32}(0)
    check(s.h, ccall((:bnl_active_bounds, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Int64}, Ref{Int32}),
                     s.h, x, st, delta, idx, cnt))
    return Int.(idx[1:cnt[]] .+ 1)
end
# add_active!(lincons, chol_aat, ind::Int) :240-249 and add_active!(lincons, chol_aat, indx::Vector{Int}) :252-261
function b200_add_active!(lincons::MixedConstraints{Float64}, chol_aat::Cholesky, indx::Vector{Int})
    s = lincons_solver(lincons)
    idx0 = Int64.(indx .- 1)                                      # 0-based across the ABI; ind = -1 -> BoundsError like :631
    check(s.h, ccall((:bnl_add_active, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int32), s.h, idx0, length(idx0)))
    pull_lincons!(s, lincons)
    return
end
b200_add_active!(lincons::MixedConstraints{Float64}, chol_aat::Cholesky, ind::Int) = b200_add_active!(lincons, chol_aat, [ind])

# the handle of a lincons used outside solve_subproblem (unit tests of the struct): M = 1 residual row, no nonlinear constraint
lincons_solver(lincons::MixedConstraints{Float64}) = haskey(SOLVERS, lincons) ?
    (s = SOLVERS[lincons]; push_fixvars!(s, lincons); s) : solver_for(lincons, 1, 0)

# inner_step(x, g, H, chol_aat, lincons, delta, nb_minor_step, kappa2, kappa3) :394-404 -> (s, model_reduction); mutates lincons
function b200_inner_step(x::Vector{Float64}, g::Vector{Float64}, H::AlHessian{Float64}, chol_aat::Cholesky,
                         lincons::MixedConstraints{Float64}, delta::Float64, nb_minor_step::Int, kappa2::Float64, kappa3::Float64)
    s = solver_for(lincons, size(H.J, 1), size(H.C, 1))
    bind_hessian!(s, H)
    set_params!(s; kappa2, kappa3, max_minor_iter=nb_minor_step)
    step = similar(x); pred = Ref{Cdouble}(0)
    check(s.h, ccall((:bnl_inner_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ref{Cdouble}),
                     s.h, x, g, delta, step, pred))
    pull_lincons!(s, lincons)                                     # lincons.fixvars / lincons.chol as inner_step leaves them (:441-452)
    return step, pred[]
end

# solve_subproblem(x0,y,mu,residuals,nlconstraints,jac_res,jac_nlcons,chol_aat,lincons,nb_minor_step,k_max,omega_tol,
#                  eta1,eta2,gamma1,gamma2,kappa2,kappa3; output_file) :303-322 -> (x, cx, pix)
function b200_solve_subproblem(x0::Vector{Float64}, y::Vector{Float64}, mu::Float64, residuals, nlconstraints, jac_res, jac_nlcons,
                               chol_aat::Cholesky, lincons::MixedConstraints{Float64}, nb_minor_step::Int, k_max::Int,
                               omega_tol::Float64, eta1::Float64, eta2::Float64, gamma1::Float64, gamma2::Float64,
                               kappa2::Float64, kappa3::Float64; output_file::IO=stdout)
    M, p = length(residuals(x0)), length(nlconstraints(x0))
    s = solver_for(lincons, M, p)
    use_callbacks!(s, residuals, jac_res, nlconstraints, jac_nlcons)
    set_params!(s; eta1, eta2, gamma1, gamma2, kappa2, kappa3, max_minor_iter=nb_minor_step, max_inner_iter=k_max)
    ccall((:bnl_reset_stats, LIB), Cint, (Ptr{Cvoid},), s.h)
    x = similar(x0); cx = Vector{Float64}(undef, p); pix = Ref{Cdouble}(Inf)
    check(s.h, ccall((:bnl_solve_subproblem, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}),
        s.h, x0, y, mu, omega_tol, x, cx, pix))
    pull_lincons!(s, lincons)                                     # criticality_measure used this active set (:369, trap T7)
    print_inner_log(s, output_file)                               # print_inner_iter lines (src/misc.jl:70-80), :356
    return x, cx, pix[]
end

struct InnerRecord
    k::Int32; nb_fix::Int32
    mx::Cdouble; norm_s::Cdouble; delta::Cdouble; rho::Cdouble; pix::Cdouble; pred::Cdouble; omega_tol::Cdouble
    breakpoints_cum::Int64; cg_cum::Int64
end
function print_inner_log(s::Solver, io::IO)
    cnt = Ref{Int32}(0)
    ccall((:bnl_get_inner_log, LIB), Cint, (Ptr{Cvoid}, Ptr{InnerRecord}, Int32, Ref{Int32}), s.h, C_NULL, 0, cnt)
    recs = Vector{InnerRecord}(undef, cnt[])
    ccall((:bnl_get_inner_log, LIB), Cint, (Ptr{Cvoid}, Ptr{InnerRecord}, Int32, Ref{Int32}), s.h, recs, cnt[], cnt)
    for r in recs
        BEnlsip.print_inner_iter(Int(r.k), r.mx, r.norm_s, r.delta, r.rho; io=io)
    end
end

"""
    enable!(; device=0)

Replace the reference's hot-path methods by the B200 ones (method overwrite on the reference's own signatures), so that
`BEnlsip.tralcnllss` and any code written against `AlHessian` / `MixedConstraints` runs on the GPU without modification.
"""
function enable!(; device::Integer=0)
    DEVICE[] = device
    @eval BEnlsip begin
        Base.:*(H::AlHessian{Float64}, v::Vector{Float64}) = $(b200_mul)(H, v)
        vthv(H::AlHessian{Float64}, v::Vector{Float64}) = $(b200_vthv)(H, v)
        projection(lincons::MixedConstraints{Float64}, r::Vector{Float64}) = $(b200_projection)(lincons, r)
        projection!(lincons::MixedConstraints{Float64}, r::Vector{Float64}, v::Vector{Float64}) = $(b200_projection!)(lincons, r, v)
        active_bounds!(lincons::MixedConstraints{Float64}, x::Vector{Float64}, chol_aat::Cholesky{Float64,Matrix{Float64}};
                       atol::Float64=sqrt(eps(Float64))) = $(b200_active_bounds!)(lincons, x, chol_aat; atol)
        active_bounds(lincons::MixedConstraints{Float64}, x::Vector{Float64}, s::Vector{Float64}, delta::Float64;
                      atol::Float64=sqrt(eps(Float64))) = $(b200_active_bounds)(lincons, x, s, delta; atol)
        add_active!(lincons::MixedConstraints{Float64}, chol_aat::Cholesky{Float64,Matrix{Float64}}, ind::Int) =
            $(b200_add_active!)(lincons, chol_aat, ind)
        add_active!(lincons::MixedConstraints{Float64}, chol_aat::Cholesky{Float64,Matrix{Float64}}, indx::Vector{Int}) =
            $(b200_add_active!)(lincons, chol_aat, indx)
        inner_step(x::Vector{Float64}, g::Vector{Float64}, H::AlHessian{Float64}, chol_aat::Cholesky{Float64,Matrix{Float64}},
                   lincons::MixedConstraints{Float64}, delta::Float64, nb_minor_step::Int, kappa2::Float64, kappa3::Float64) =
            $(b200_inner_step)(x, g, H, chol_aat, lincons, delta, nb_minor_step, kappa2, kappa3)
        solve_subproblem(x0::Vector{Float64}, y::Vector{Float64}, mu::Float64, residuals::F1, nlconstraints::F2, jac_res::F3,
                         jac_nlcons::F4, chol_aat::Cholesky{Float64,Matrix{Float64}}, lincons::MixedConstraints{Float64},
                         nb_minor_step::Int, k_max::Int, omega_tol::Float64, eta1::Float64, eta2::Float64, gamma1::Float64,
                         gamma2::Float64, kappa2::Float64, kappa3::Float64; output_file::IO=stdout) where
                         {F1<:Function,F2<:Function,F3<:Function,F4<:Function} =
            $(b200_solve_subproblem)(x0, y, mu, residuals, nlconstraints, jac_res, jac_nlcons, chol_aat, lincons, nb_minor_step,
                                     k_max, omega_tol, eta1, eta2, gamma1, gamma2, kappa2, kappa3; output_file)
    end
    return nothing
end

# Convenience: the package's entry point with the hot path on the GPU (same signature and keyword arguments, :167-197).
function tralcnllss(args...; device::Integer=0, kwargs...)
    enable!(; device)
    return BEnlsip.tralcnllss(args...; kwargs...)
end

end # module
