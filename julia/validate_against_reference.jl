# validate_against_reference.jl -- turns "trajectory parity pinned by the restatement only" (DESIGN.md section 5) into a check
# against the real reference, on a machine that has Julia, the BEnlsip.jl package and a B200 with libbenlsip_b200.so built.
#
#     julia --project=/path/to/BEnlsip.jl julia/validate_against_reference.jl [M n m_lin]
#
# It solves one synthetic problem (box bounds + m_lin linear equalities + one sphere constraint, defined below -- not taken from
# the reference) twice through the package's own entry point `BEnlsip.tralcnllss`: first with the package untouched, then after
# `BEnlsipB200.enable!()` has put the B200 methods over the package's hot path.  Both runs write the reference's log
# (`../test/benlsip.out`, src/basic_tralcnlss.jl:4,202); the logs are compared line by line -- same text, every printed number
# equal up to one unit of its last printed digit -- up to the first inner iteration whose rho is a ratio of rounding noise
# (|ared| <= 64 eps |mx|, the criterion of tests/parity.py; later lines are not compared), and the final iterates are reported.
#
# NOT EXECUTED in this repository (no Julia in the build image or on the GPU box); the same comparison is run against the NumPy
# restatement of the reference by tests/test_gpu_round2.py::test_native_log_equals_the_oracle_log.
using LinearAlgebra, Random, Printf
using BEnlsip
include(joinpath(@__DIR__, "BEnlsipB200.jl"))

M = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 600
n = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 24
m_lin = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 4

# ---- the problem: r(x) = tanh.(D x) - d, sphere constraint x'x = rho2, A x = b, -1 <= x <= 1 ------------------------------------
rng = MersenneTwister(3)
D = randn(rng, M, n) ./ sqrt(n)
x_star = clamp.(0.6 .* randn(rng, n), -0.9, 0.9)
d = tanh.(D * x_star) .+ 1e-3 .* randn(rng, M)
A = randn(rng, m_lin, n)
b = A * x_star
rho2 = dot(x_star, x_star)
x_l, x_u = fill(-1.0, n), fill(1.0, n)
x0 = clamp.(x_star .+ 0.3 .* randn(rng, n), -1.0, 1.0)

residuals(x) = tanh.(D * x) .- d
jac_res(x) = (1 .- tanh.(D * x) .^ 2) .* D
nlconstraints(x) = [dot(x, x) - rho2]
jac_nlcons(x) = reshape(2 .* x, 1, n)

# the package writes "../test/benlsip.out" relative to the working directory: run inside <tmp>/test and read the file back
function solve_and_grab_log()
    root = mktempdir()
    wd = joinpath(root, "test")
    mkpath(wd)
    local out
    cd(wd) do
        out = BEnlsip.tralcnllss(x0, residuals, jac_res, nlconstraints, jac_nlcons, A, b, x_l, x_u;
                                 max_outer_iter = 60, max_inner_iter = 200)
    end
    return out, readlines(joinpath(wd, "benlsip.out"))
end

const NUM = r"[-+]?(?:\d+\.\d*(?:[eE][-+]?\d+)?|NaN|Inf)"
# two printed numbers agree when they differ by at most one unit of the last printed digit of the coarser one
function same_number(a::AbstractString, b::AbstractString)
    a == b && return true
    (occursin("NaN", a) || occursin("NaN", b) || occursin("Inf", a) || occursin("Inf", b)) && return false
    va, vb = parse(Float64, a), parse(Float64, b)
    digits_after(s) = (m = match(r"\.(\d*)", s); m === nothing ? 0 : length(m.captures[1]))
    expo(s) = (m = match(r"[eE]([-+]?\d+)", s); m === nothing ? 0 : parse(Int, m.captures[1]))
    ulp = max(10.0^(expo(a) - digits_after(a)), 10.0^(expo(b) - digits_after(b)))
    return abs(va - vb) <= 1.0000001 * ulp
end
function same_line(a::AbstractString, b::AbstractString; numbers::Bool = true)
    replace(a, NUM => "#") == replace(b, NUM => "#") || return false
    numbers || return true
    na, nb = collect(eachmatch(NUM, a)), collect(eachmatch(NUM, b))
    return all(same_number(x.match, y.match) for (x, y) in zip(na, nb))
end

# index of the first inner-iteration line ("iter  AL value  ||s||  Delta  rho") whose rho is a ratio of rounding noise.  The log
# does not print ared or pred, so this is a heuristic on what it does print: an ACCEPTED step (rho > eta1 = 0.25) after which the
# AL value moved by no more than 64 eps |mx| (the criterion of tests/parity.py), or a rho no meaningful Gauss-Newton step produces
# (rho >= 2.5 or rho <= -10: the tails of the committed goldens show 2.98, 26.1, -63.3, -4261 ...).
function first_noise_line(lines)
    prev = nothing                       # (mx, rho) of the previous inner line of the same subproblem
    for (k, ln) in enumerate(lines)
        m = match(r"^\s*(\d+)\s+(\S+)\s+(\S+)\s+(\S+)\s+(\S+)\s*$", ln)
        mx = m === nothing ? nothing : tryparse(Float64, m.captures[2])
        rho = m === nothing ? nothing : tryparse(Float64, m.captures[5])
        if mx === nothing || rho === nothing
            prev = nothing
            continue
        end
        if prev !== nothing && parse(Int, m.captures[1]) > 1 && prev[2] > 0.25 && abs(mx - prev[1]) <= 64 * eps() * abs(mx)
            return k - 1                 # the previous line's accepted step changed the AL value by rounding noise only
        end
        (isfinite(rho) && (rho >= 2.5 || rho <= -10.0)) && return k
        prev = (mx, rho)
    end
    return length(lines) + 1
end

(x_ref, y_ref), log_ref = solve_and_grab_log()
BEnlsipB200.enable!()
(x_gpu, y_gpu), log_gpu = solve_and_grab_log()

strict_until = first_noise_line(log_ref)          # from this line on the reference's own decisions are rounding noise: not compared
nbad = 0
for k in 1:min(length(log_ref), length(log_gpu), strict_until - 1)
    if !same_line(log_gpu[k], log_ref[k])
        global nbad += 1
        nbad <= 5 && @printf("line %d differs\n  reference: %s\n  b200     : %s\n", k, log_ref[k], log_gpu[k])
    end
end
short = length(log_gpu) < min(length(log_ref), strict_until - 1)
@printf("log lines: reference %d, b200 %d; compared up to line %d; %d differing lines%s\n",
        length(log_ref), length(log_gpu), min(length(log_ref), strict_until - 1), nbad, short ? "; b200 log is SHORTER" : "")
@printf("final x: relative difference %.3e;  y: reference %s, b200 %s\n", norm(x_gpu - x_ref) / norm(x_ref), string(y_ref), string(y_gpu))
@printf("feasibility of the b200 iterate: |c(x)| = %.2e, |Ax - b|_inf = %.2e, bounds %s\n", abs(nlconstraints(x_gpu)[1]),
        norm(A * x_gpu - b, Inf), all(x_l .<= x_gpu .<= x_u) ? "ok" : "VIOLATED")
exit((nbad == 0 && !short) ? 0 : 1)
