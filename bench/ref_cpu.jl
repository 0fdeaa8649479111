# bench/ref_cpu.jl -- the REFERENCE's own CPU path for the hot path of this repository, timed with the reference's own code
# (NEP-PACK / NonlinearEigenproblems.jl v1.1.1 on SparseArrays + UMFPACK + OpenBLAS).  Julia is not part of the build or GPU
# images, so this script has never been executed here (stated in BASELINE.md section 3 and DESIGN.md section 6); it is the
# recipe for a box that has Julia:
#
#     julia --project=<NEP-PACK checkout> -p <host cores> bench/ref_cpu.jl [steps]
#
# Same inputs as bench.py (SURVEY.md 8(d)): gun from the reference's own text files, probe = MSWS 1-2u stream, contour
# sigma = 150^2, radius = 500, N = 128, k = 20; C4 = synthetic degree-3 PEP on the 1000 x 1000 grid built by the same generator
# (read from the .npz files that `python tools/export_c4.py` writes, if present).  Prints one JSON line per measurement in the
# unit bench.py uses (quadrature-point solves/s, GB/s of algorithmic bytes).
using Distributed
@everywhere using NonlinearEigenproblems, LinearAlgebra, SparseArrays, Random
using Printf

steps = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 3
cores = max(nworkers(), 1)
cpu = Sys.cpu_info()[1].model

# ---- C3: contour_beyn on gun -------------------------------------------------------------------------------------------
@everywhere const GUN = nep_gallery("nlevp_native_gun")
n = size(GUN, 1); k = 20; N = 128; sigma = 150.0^2; radius = 500.0

# MSWS probe, bit-identical to bench.py's (basic_random_examples.jl:73-105 is the generator both sides use)
function msws_probe(n, k)
    rng = NonlinearEigenproblems.Gallery.MSWS_RNG()
    V = Matrix{ComplexF64}(undef, n, k)
    for j = 1:k, i = 1:n
        V[i, j] = 1 - 2 * NonlinearEigenproblems.Gallery.gen_rng_float(rng)
    end
    return V
end
Vh = msws_probe(n, k)

# (a) the reference's contour_beyn as it ships: serial trapezoidal rule, one UMFPACK factorisation + k solves per node
function time_contour(creator)
    contour_beyn(GUN; σ=sigma, radius=radius, N=8, k=k, neigs=1, linsolvercreator=creator, sanity_check=false)   # compile
    t = @elapsed for _ = 1:steps
        contour_beyn(GUN; σ=sigma, radius=radius, N=N, k=k, neigs=1, linsolvercreator=creator, sanity_check=false)
    end
    return N * steps / t
end
for (name, creator) in (("FactorizeLinSolverCreator", FactorizeLinSolverCreator()), ("BackslashLinSolverCreator", BackslashLinSolverCreator()))
    v = time_contour(creator)
    @printf("{\"impl\": \"reference-julia\", \"metric\": \"contour_beyn quadrature-point solves/sec (gun, N=128, k=20)\", \"value\": %.3f, \"unit\": \"solves/s\", \"cores\": 1, \"linsolver\": \"%s\", \"cpu\": \"%s\"}\n", v, name, cpu)
end

# (b) the docs' @distributed quadrature (docs/src/tutorial_contour.md: each worker owns a copy of the NEP): nodes over all workers
@everywhere function node_block(lam, w0, w1, Vh)
    F = lu(compute_Mder(GUN, lam))
    X = F \ Vh
    return (w0 .* X, w1 .* X)
end
function distributed_moments()
    h = 2π / N; t = h .* (0:N-1)
    g = radius .* cis.(t); gp = im .* g
    parts = pmap(i -> node_block(g[i] + sigma, gp[i] * h, gp[i] * g[i] * h, Vh), 1:N)
    return sum(p[1] for p in parts), sum(p[2] for p in parts)
end
distributed_moments()
t = @elapsed for _ = 1:steps; distributed_moments(); end
@printf("{\"impl\": \"reference-julia\", \"metric\": \"contour_beyn quadrature-point solves/sec (gun, N=128, k=20)\", \"value\": %.3f, \"unit\": \"solves/s\", \"cores\": %d, \"linsolver\": \"lu(compute_Mder) on every worker (pmap over the nodes)\", \"cpu\": \"%s\"}\n", N * steps / t, cores, cpu)

# ---- C4: compute_MM / compute_Mlincomb of the synthetic PEP (needs the exported matrices) ---------------------------------
c4 = joinpath(@__DIR__, "c4")
if isdir(c4)
    using DelimitedFiles
    function read_coo(f)
        d = readdlm(f); sparse(Int.(d[:, 1]), Int.(d[:, 2]), d[:, 3])
    end
    Av = [read_coo(joinpath(c4, "A$(i).txt")) for i = 0:3]
    pep = PEP(Av)
    nn = size(pep, 1); nnzu = nnz(Av[1]); lam = 0.3 + 0.2im
    for kk in (1, 8, 20)
        V = Matrix{ComplexF64}(undef, nn, kk); rng = MersenneTwister(0); V .= 1 .- 2 .* rand(rng, nn, kk)
        S = Matrix{ComplexF64}(lam * I, kk, kk)
        compute_MM(pep, S, V)
        t = @elapsed for _ = 1:steps; compute_MM(pep, S, V); end
        bytes = nnzu * (4 * 8 + 4) + 4 * (nn + 1) + 2 * 16 * nn * kk          # SURVEY 8(d)
        @printf("{\"impl\": \"reference-julia\", \"metric\": \"compute_MM algorithmic GB/s (C4, k=%d)\", \"value\": %.3f, \"unit\": \"GB/s\", \"cores\": 1, \"cpu\": \"%s\"}\n", kk, bytes * steps / t / 1e9, cpu)
    end
end

# ---- C2: iar on gun (the representable variant of the survey's configuration, DESIGN.md section 5) ---------------------------
v0 = ones(ComplexF64, n)
iar(GUN; σ=250.0^2, γ=1000.0, maxit=10, neigs=Inf, v=v0, tol=1e-10, check_error_every=10)
t = @elapsed iar(GUN; σ=250.0^2, γ=1000.0, maxit=100, neigs=Inf, v=v0, tol=1e-10, check_error_every=10)
@printf("{\"impl\": \"reference-julia\", \"metric\": \"iar on gun, m=100, gamma=1000\", \"value\": %.3f, \"unit\": \"s\", \"cores\": 1, \"cpu\": \"%s\"}\n", t, cpu)
