# bench_reference.jl -- the UNMODIFIED reference (AdaProx.jl) on the inputs the device run used.
#
# NOT EXECUTED: Julia is not installed in the build image (SURVEY.md 8d, item ii).  It exists so that whoever has Julia
# can put the reference's own numbers next to bench.py's:
#
#   python tools/dump_reference_inputs.py --m 400 --n 1000 --out /tmp/c1        # same counter-RNG instance as the device run
#   julia --project=/path/to/adaptive-proximal-algorithms julia/bench_reference.jl /tmp/c1 [tol] [maxit]
#
# It reads A (column-major Float64, the layout dump_reference_inputs.py writes), b and gamma0 = 1/Lf, runs
# AdaProx.adaptive_proxgrad (src/AdaProx.jl:418) with the experiment's own oracle struct (experiments/lasso/runme.jl:16-27)
# and prints ONE JSON line shaped like bench.py's reference arm: iterations per second, the thread counts actually
# used, and the stepsize / objective prefix for a parity check against the device records.
using LinearAlgebra
using Printf
using AdaProx
using ProximalOperators: NormL1

struct LinearLeastSquares{TA,Tb}          # experiments/lasso/runme.jl:16-19
    A::TA
    b::Tb
end
(f::LinearLeastSquares)(w) = 0.5 * norm(f.A * w - f.b)^2
function AdaProx.eval_with_pullback(f::LinearLeastSquares, w)   # experiments/lasso/runme.jl:21-27
    res = f.A * w - f.b
    linear_least_squares_pullback() = f.A' * res
    return 0.5 * norm(res)^2, linear_least_squares_pullback
end

function read_f64(path, dims...)
    v = Array{Float64}(undef, dims...)
    open(io -> read!(io, v), path)
    return v
end

function main()
    dir = ARGS[1]
    tol = length(ARGS) >= 2 ? parse(Float64, ARGS[2]) : 1e-6
    maxit = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 10_000
    meta = Dict(split(l, "=")[1] => split(l, "=")[2] for l in eachline(joinpath(dir, "meta.txt")))
    m, n = parse(Int, meta["m"]), parse(Int, meta["n"])
    lam, gamma0 = parse(Float64, meta["lambda"]), parse(Float64, meta["gamma0"])
    A = read_f64(joinpath(dir, "A.f64"), m, n)          # column-major, as Julia stores it
    b = read_f64(joinpath(dir, "b.f64"), m)
    f, g = LinearLeastSquares(A, b), NormL1(lam)
    rule = AdaProx.OurRule(gamma = gamma0)
    AdaProx.adaptive_proxgrad(zeros(n); f = f, g = g, rule = rule, tol = tol, maxit = 5)     # compile
    t0 = time_ns()
    sol, numit = AdaProx.adaptive_proxgrad(zeros(n); f = AdaProx.Counting(f), g = g, rule = rule, tol = tol, maxit = maxit)
    secs = (time_ns() - t0) * 1e-9
    obj = f(sol) + g(sol)
    @printf("{\"impl\": \"reference\", \"kind\": \"reference\", \"metric\": \"AdaPGM iters/sec on %dx%d fp64 lasso\", \"value\": %.6f, \"unit\": \"it/s\", ", m, n, numit / secs)
    @printf("\"iterations\": %d, \"seconds\": %.6f, \"objective\": %.17g, \"julia_threads\": %d, \"blas_threads\": %d, \"blas\": \"%s\"}\n",
            numit, secs, obj, Threads.nthreads(), BLAS.get_num_threads(), string(BLAS.get_config()))
end

main()
