# BEnlsipB200.jl -- Julia shim: keeps BEnlsip.jl's entry points and structs for the inner Gauss-Newton
# trust-region solve and forwards them to libbenlsip_b200.so (include/benlsip_b200.h) through `ccall`.
# NOT EXECUTED in this repository's CI (no Julia in the build image); it is the literal transliteration of
# benlsip.jl_b200/__init__.py, which is the tested host side.  Usage (see INTEGRATION.md):
#
#     using BEnlsip; include("BEnlsipB200.jl"); using .BEnlsipB200
#     x, y = BEnlsipB200.tralcnllss(x0, r, jac_r, c, jac_c, A, b, x_l, x_u)      # same signature / kwargs
#
module BEnlsipB200

using LinearAlgebra

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

mutable struct Solver
    h::Ptr{Cvoid}
    n::Int; M::Int; p::Int
    cbs::Any                     # keeps the @cfunction closures alive
    function Solver(device::Integer=0)
        hp = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:bnl_create, LIB), Cint, (Cint, Ptr{Ptr{Cvoid}}), device, hp)
        rc == 0 || error("bnl_create: " * unsafe_string(ccall((:bnl_status_string, LIB), Cstring, (Cint,), rc)))
        s = new(hp[], 0, 0, 0, nothing)
        finalizer(s -> ccall((:bnl_destroy, LIB), Cvoid, (Ptr{Cvoid},), s.h), s)
        return s
    end
end

# MixedConstraints(A, cholesky(A*A'); l, u)   src/polyhedral_constraints.jl:9-18, src/basic_tralcnlss.jl:206
function set_problem!(s::Solver, M::Integer, A::Matrix{Float64}, x_l::Vector{Float64}, x_u::Vector{Float64}, p::Integer)
    m, n = size(A)
    GC.@preserve A x_l x_u check(s.h, ccall((:bnl_set_problem, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Int64, Int32, Int32, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        s.h, M, M, 0, n, m, p, m == 0 ? C_NULL : pointer(A), pointer(x_l), pointer(x_u)))
    s.n, s.M, s.p = n, M, p
end

# the four closures of tralcnllss (src/basic_tralcnlss.jl:167-176) as C callbacks; matrices cross column-major
function use_callbacks!(s::Solver, residuals, jac_res, nlconstraints, jac_nlcons)
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
    check(s.h, ccall((:bnl_use_callbacks, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                     s.h, cbs[1], cbs[2], cbs[3], cbs[4], C_NULL))
end

function set_params!(s::Solver; eta1, eta2, gamma1, gamma2, kappa2, kappa3, max_minor_iter, max_inner_iter)
    p = Ref(BnlParams(eta1, eta2, gamma1, gamma2, kappa2, kappa3, 0.1, sqrt(eps()), sqrt(eps()), 1e-10,
                      max_minor_iter, max_inner_iter))
    check(s.h, ccall((:bnl_set_params, LIB), Cint, (Ptr{Cvoid}, Ref{BnlParams}), s.h, p))
end

# opt-in modes (INTEGRATION.md 2b); defaults are the reference's semantics
const HESSIAN_MATRIX_FREE, HESSIAN_GRAM = Int32(0), Int32(1)
const CAUCHY_LITERAL, CAUCHY_INCREMENTAL = Int32(0), Int32(1)
set_hessian_mode!(s::Solver, mode::Int32) = check(s.h, ccall((:bnl_set_hessian_mode, LIB), Cint, (Ptr{Cvoid}, Int32), s.h, mode))
set_cauchy_mode!(s::Solver, mode::Int32) = check(s.h, ccall((:bnl_set_cauchy_mode, LIB), Cint, (Ptr{Cvoid}, Int32), s.h, mode))

# Base.:*(H::AlHessian, v) :102-106 and vthv :92-96 on the (J, C, mu) the handle currently holds
hess_mul(s::Solver, v::Vector{Float64}) = (out = similar(v);
    check(s.h, ccall((:bnl_hess_mul, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.h, v, out)); out)
vthv(s::Solver, v::Vector{Float64}) = (out = Ref{Cdouble}(0);
    check(s.h, ccall((:bnl_vthv, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ref{Cdouble}), s.h, v, out)); out[])
# projection(lincons, r)  src/polyhedral_constraints.jl:150-170
projection(s::Solver, r::Vector{Float64}) = (out = similar(r);
    check(s.h, ccall((:bnl_project, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.h, r, out)); out)
# lincons.fixvars as a BitVector (chunks cross the ABI unchanged)
function fixvars(s::Solver)
    b = BitVector(undef, s.n)
    cnt = Ref{Int32}(0)
    check(s.h, ccall((:bnl_get_fixvars, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Ref{Int32}), s.h, b.chunks, cnt))
    return b
end
function set_fixvars!(s::Solver, b::BitVector)
    check(s.h, ccall((:bnl_set_fixvars, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt64}), s.h, b.chunks))
end

# inner_step(x,g,H,chol_aat,lincons,delta,nb_minor_step,kappa2,kappa3) :394-460 -> (s, model_reduction)
function inner_step(sv::Solver, x::Vector{Float64}, g::Vector{Float64}, delta::Float64)
    s = similar(x); pred = Ref{Cdouble}(0)
    check(sv.h, ccall((:bnl_inner_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ref{Cdouble}),
                      sv.h, x, g, delta, s, pred))
    return s, pred[]
end

# solve_subproblem(x0,y,mu,...,omega_tol,...) :303-378 -> (x, cx, pix)
function solve_subproblem(s::Solver, x0::Vector{Float64}, y::Vector{Float64}, mu::Float64, omega_tol::Float64)
    x = similar(x0); cx = Vector{Float64}(undef, s.p); pix = Ref{Cdouble}(Inf)
    check(s.h, ccall((:bnl_solve_subproblem, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}),
        s.h, x0, y, mu, omega_tol, x, cx, pix))
    return x, cx, pix[]
end

# tralcnllss :167-298 -- outer loop stays in Julia; only solve_subproblem (and J'r for the initial multipliers) cross the ABI
function tralcnllss(x0::Vector{T}, residuals, jac_res, nlconstraints, jac_nlcons, A::Matrix{T}, b::Vector{T},
        x_l::Vector{T}, x_u::Vector{T};
        mu0::T=T(10), tau::T=T(100), omega0::T=T(1), eta0::T=T(1), feas_tol::T=sqrt(eps(T)), crit_tol::T=sqrt(eps(T)),
        k_crit::T=T(1), k_feas::T=T(0.1), beta_crit::T=T(1), beta_feas::T=T(0.9), eta1::T=T(0.25), eta2::T=T(0.75),
        gamma1::T=T(0.0625), gamma2::T=T(2), gamma_c::T=T(10), kappa1::T=T(1e-2), kappa2::T=T(0.1), kappa3::T=T(0.1),
        max_outer_iter::Int=500, max_inner_iter::Int=500, max_minor_iter::Int=50, device::Int=0) where {T<:Float64}
    @assert (0 < eta1 <= eta2 < 1) && (0 < gamma1 < 1 < gamma2) "Invalid trust region updates paramaters"
    n = length(x0)
    x = copy(x0)
    rx = residuals(x); cx = nlconstraints(x)
    s = Solver(device)
    set_problem!(s, length(rx), A, x_l, x_u, length(cx))
    use_callbacks!(s, residuals, jac_res, nlconstraints, jac_nlcons)
    set_params!(s; eta1, eta2, gamma1, gamma2, kappa2, kappa3, max_minor_iter, max_inner_iter)
    mu = mu0
    omega, eta = omega0 / (mu0^k_crit), eta0 / (mu0^k_feas)
    # least_squares_multipliers :887-903
    y = if length(cx) > 0
        check(s.h, ccall((:bnl_eval_jacobian, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.h, x))
        g = similar(x)
        check(s.h, ccall((:bnl_jtw, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.h, rx, g))
        C = jac_nlcons(x); ch = cholesky(C * C'); ch.U \ (ch.L \ (-C * g))
    else
        T[]
    end
    set_fixvars!(s, falses(n))
    first_order_critical = false
    outer_iter = 1
    while !first_order_critical && outer_iter <= max_outer_iter
        x_next, cx_next, pix = solve_subproblem(s, x, y, mu, omega)          # <-- the C-ABI boundary (:249-268)
        feas_measure = norm(cx_next)
        if feas_measure <= eta
            x .= x_next; cx = cx_next
            first_order_critical = pix <= crit_tol && feas_measure <= feas_tol
            if !first_order_critical
                y = y + mu * cx
                omega /= mu^beta_crit
                eta /= mu^beta_feas
            end
        else
            mu *= tau
            omega = omega0 / (mu^k_crit)
            eta = eta0 / (mu^k_feas)
        end
        outer_iter += 1
    end
    return x, y
end

end # module
