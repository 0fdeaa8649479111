# NEPB200.jl -- Julia shim that drops libnepb200.so (B200 / sm_100a) behind NEP-PACK's own plugin interfaces.
#
# Nothing in NEP-PACK's solver loops changes: the new path is selected only by the *types* passed in
# (`B200SPMF` as the NEP, `B200LinSolverCreator` as the linsolvercreator, `B200Trapezoidal` as the integrator),
# exactly the extension route the reference documents (src/LinSolvers.jl:64-91, docs/src/tutorial_linsolve.md:108-187,
# src/gallery_extra/waveguide/Waveguide.jl:324,504,554).  Each method below cites the reference method it replaces.
#
#   using NonlinearEigenproblems; include("NEPB200.jl"); using .NEPB200
#   nep  = nep_gallery("nlevp_native_gun")
#   dnep = B200SPMF(nep)                                   # union-pattern CSR of K, -M, W1, W2 resident in HBM
#   λ,v  = iar(dnep; σ=250.0^2, γ=300.0^2-200.0^2, maxit=100, neigs=Inf,
#              linsolvercreator=B200LinSolverCreator())    # compute_Mlincomb! and lin_solve run on the GPU
#   λ,v  = contour_beyn(dnep; σ=150.0^2, radius=500, N=128, k=20)   # all 128 nodes in one batched device call
#
# This file cannot be executed in the build container (no Julia there); the C ABI it binds is exercised through the
# ctypes twin in nonlineareigenproblems.jl_b200/_lib.py, signature by signature (tests/test_abi.py).
module NEPB200

using LinearAlgebra, SparseArrays, Libdl
using NonlinearEigenproblems
using NonlinearEigenproblems.NEPCore, NonlinearEigenproblems.NEPTypes, NonlinearEigenproblems.LinSolvers
import NonlinearEigenproblems.NEPCore: compute_Mder, compute_Mlincomb, compute_Mlincomb!, compute_MM, compute_resnorm
import NonlinearEigenproblems.NEPTypes: get_Av, get_fv
import NonlinearEigenproblems.LinSolvers: lin_solve, create_linsolver
import NonlinearEigenproblems.NEPSolver: contour_beyn, integrate_interval, MatrixIntegrator
import Base: size
import SparseArrays: issparse

export B200SPMF, B200LinSolver, B200LinSolverCreator, B200BackslashLinSolverCreator, B200Trapezoidal,
       b200_residual_norms, b200_comm_unique_id, b200_comm_init

const CF = ComplexF64
const libpath = Ref(joinpath(@__DIR__, "..", "libnepb200.so"))
const lib = Ref{Ptr{Cvoid}}(C_NULL)
sym(s::Symbol) = (lib[] == C_NULL && (lib[] = Libdl.dlopen(libpath[])); Libdl.dlsym(lib[], s))

const COEF_SCALAR, COEF_DIAG, COEF_GENERAL = Cint(0), Cint(1), Cint(2)

# every entry point returns 0 or a negative status; the message of the last failure is kept per thread
function chk(status::Cint)
    status == 0 && return
    msg = unsafe_string(ccall(sym(:nepb_last_error), Cstring, ()))
    status == -3 && throw(LinearAlgebra.SingularException(0))   # zero / non-finite pivot, like lu(M) in the reference
    error("libnepb200 status $status: $msg")
end

# ------------------------------------------------------------------------------------------------
# a1: the SPMF operator (src/NEPTypes.jl:162-237; union pattern = form_aligned_sparsity_patterns :244-274)
# ------------------------------------------------------------------------------------------------
mutable struct B200SPMF <: AbstractSPMF{CF}
    n::Int
    A::Vector{SparseMatrixCSC}      # kept for get_Av (errmeasure.jl:177-180, rk_nep.jl:102-110 read them)
    fi::Vector{Function}
    h::Ptr{Cvoid}
    nnz_union::Int
    function B200SPMF(A::Vector{<:SparseMatrixCSC}, fi::Vector)
        length(A) == length(fi) || error("Inconsistency: Number of supplied matrices = $(length(A)) but the number of supplied functions are = $(length(fi))")
        n = size(A[1], 1)
        all(a -> size(a) == (n, n), A) || error("The dimensions of the matrices mismatch")
        cplx = any(a -> eltype(a) <: Complex, A)
        T = cplx ? CF : Float64
        AA = [SparseMatrixCSC{T,Int64}(a) for a in A]
        colptr = [pointer(a.colptr) for a in AA]; rowval = [pointer(a.rowval) for a in AA]; nzval = [Ptr{Cvoid}(pointer(a.nzval)) for a in AA]
        out = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve AA chk(ccall(sym(:nepb_spmf_create), Cint,
            (Int64, Cint, Ptr{Ptr{Int64}}, Ptr{Ptr{Int64}}, Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{Ptr{Cvoid}}),
            n, length(AA), colptr, rowval, nzval, cplx ? 1 : 0, 1, out))       # index_base = 1: Julia's CSC as is
        nnzu = Ref{Int64}(0)
        chk(ccall(sym(:nepb_spmf_info), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Cint}, Ptr{Int64}, Ptr{Cint}), out[], C_NULL, C_NULL, nnzu, C_NULL))
        nep = new(n, Vector{SparseMatrixCSC}(A), Vector{Function}(fi), out[], nnzu[])
        finalizer(x -> (x.h != C_NULL && ccall(sym(:nepb_spmf_destroy), Cint, (Ptr{Cvoid},), x.h); x.h = C_NULL), nep)
        return nep
    end
end
# any AbstractSPMF with sparse terms: SPMF_NEP, PEP, DEP, SumNEP (gun = SumNEP(PEP, SPMF_NEP), NLEVP_native.jl:4-18)
B200SPMF(nep::AbstractSPMF) = B200SPMF([sparse(a) for a in get_Av(nep)], get_fv(nep))

size(nep::B200SPMF) = (nep.n, nep.n)
size(nep::B200SPMF, dim) = nep.n
issparse(nep::B200SPMF) = true
get_Av(nep::B200SPMF) = nep.A
get_fv(nep::B200SPMF) = nep.fi

coefficients(nep::B200SPMF, λ::Number) = CF[f(reshape([CF(λ)], 1, 1))[1] for f in nep.fi]
function coefficients(nep::B200SPMF, λ::Number, der::Integer)
    der == 0 && return coefficients(nep, λ)
    S = diagm(0 => fill(CF(λ), der + 1), -1 => CF.(1:der))        # Jordan trick, NEPTypes.jl:376-386
    return CF[f(S)[end, 1] for f in nep.fi]
end

# a5: compute_Mder (NEPTypes.jl:336-367): values on the union pattern, pattern fetched once
function pattern(nep::B200SPMF)
    colptr = Vector{Int64}(undef, nep.n + 1); rowval = Vector{Int64}(undef, nep.nnz_union)
    chk(ccall(sym(:nepb_spmf_pattern), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), nep.h, colptr, rowval))
    return colptr, rowval
end
function compute_Mder(nep::B200SPMF, λ::Number, i::Integer=0)
    c = coefficients(nep, λ, i)
    colptr, rowval = pattern(nep)
    nz = Vector{CF}(undef, nep.nnz_union)
    chk(ccall(sym(:nepb_spmf_mder), Cint, (Ptr{Cvoid}, Ptr{CF}, Ptr{CF}), nep.h, c, nz))
    return SparseMatrixCSC(nep.n, nep.n, colptr, rowval, nz)
end

# raw kernel call: Z = sum_i A_i (V C_i)
function apply(nep::B200SPMF, mode::Cint, V::StridedMatrix{CF}, C::Array{CF}, q::Integer)
    n, k = size(V)
    n == nep.n || error("V has $n rows, the NEP has size $(nep.n)")
    Z = Matrix{CF}(undef, n, q)
    chk(ccall(sym(:nepb_spmf_apply), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{CF}, Int64, Ptr{CF}, Ptr{CF}, Int64),
              nep.h, mode, k, q, V, stride(V, 2), C, Z, n))
    return Z
end

# a2: compute_Mlincomb! (NEPTypes.jl:972-1011 incl. the zero-entry convention :982-983; NEPCore.jl:113-160)
function compute_Mlincomb!(nep::B200SPMF, λ::Number, V::AbstractVecOrMat, a::Vector=ones(CF, size(V, 2)))
    Vm = Matrix{CF}(reshape(V, nep.n, :)); k = size(Vm, 2)
    aa = CF.(a); z = aa .== 0; aa[z] .= 1
    p = length(nep.fi)
    C = Matrix{CF}(undef, k, p)                                    # block i (k x 1) = a_1 * f_i(S)[:,1]
    if k == 1
        C[1, :] = aa[1] .* coefficients(nep, λ)
    else
        S = diagm(0 => fill(CF(λ), k), -1 => (aa[2:k] ./ aa[1:k-1]) .* (1:k-1))
        for i = 1:p; C[:, i] = aa[1] .* nep.fi[i](S)[:, 1]; end
    end
    C[z, :] .= 0                                                   # zeroed columns of V == zero coefficients
    return vec(apply(nep, COEF_GENERAL, Vm, C, 1))
end
compute_Mlincomb(nep::B200SPMF, λ::Number, V::AbstractVecOrMat) = compute_Mlincomb!(nep, λ, V)          # inputs are never modified
compute_Mlincomb(nep::B200SPMF, λ::Number, V::AbstractVecOrMat, a::Vector) = compute_Mlincomb!(nep, λ, V, a)

# a4: compute_MM (NEPTypes.jl:276-319, diagonal fast paths :299-311)
function compute_MM(nep::B200SPMF, S::AbstractMatrix, V::AbstractMatrix)
    Vm = Matrix{CF}(V); q = size(S, 1); p = length(nep.fi)
    if isdiag(S)
        d = CF.(diag(S))
        Cd = CF[nep.fi[i](reshape([d[s]], 1, 1))[1] for i = 1:p, s = 1:q]     # p x q, column-major = C[i + p*s]
        all(d .== d[1]) && return apply(nep, COEF_SCALAR, Vm, Cd[:, 1], q)
        return apply(nep, COEF_DIAG, Vm, Cd, q)
    end
    C = Array{CF}(undef, q, q, p)
    for i = 1:p; C[:, :, i] = nep.fi[i](Matrix{CF}(S)); end
    return apply(nep, COEF_GENERAL, Vm, C, q)
end

# a13: all Ritz residuals in one multi-lambda SpMM (replaces k estimate_error calls, method_iar.jl:134-135)
function b200_residual_norms(nep::B200SPMF, λv::Vector, V::AbstractMatrix)
    R = compute_MM(nep, Diagonal(CF.(λv)), V)
    return [norm(R[:, s]) / norm(V[:, s]) for s = 1:length(λv)]
end

# ------------------------------------------------------------------------------------------------
# a6, a7: LinSolver / LinSolverCreator (LinSolvers.jl:100-159, LinSolverCreators.jl:11-122)
# ------------------------------------------------------------------------------------------------
mutable struct B200LinSolver <: LinSolver
    nep::B200SPMF
    h::Ptr{Cvoid}
    refinements::Int
    function B200LinSolver(nep::B200SPMF, λ::Number, umfpack_refinements::Integer=10)
        c = coefficients(nep, λ)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        chk(ccall(sym(:nepb_lu_create), Cint, (Ptr{Cvoid}, Cint, Ptr{CF}, Ptr{Ptr{Cvoid}}), nep.h, 1, c, out))
        s = new(nep, out[], umfpack_refinements)
        finalizer(x -> (x.h != C_NULL && ccall(sym(:nepb_lu_destroy), Cint, (Ptr{Cvoid},), x.h); x.h = C_NULL), s)
        return s
    end
end
# lin_solve(solver, b; tol) returns a new array of b's shape (LinSolvers.jl:135-137); vector or n x k matrix RHS
function lin_solve(solver::B200LinSolver, b::AbstractVecOrMat; tol=0)
    B = Matrix{CF}(reshape(b, solver.nep.n, :)); X = similar(B)
    chk(ccall(sym(:nepb_lu_solve), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{CF}, Int64, Ptr{CF}, Int64, Cint, Ptr{Cdouble}),
              solver.h, 0, size(B, 2), B, size(B, 1), X, size(X, 1), solver.refinements, C_NULL))
    return b isa AbstractVector ? vec(X) : X
end

struct B200LinSolverCreator <: LinSolverCreator
    umfpack_refinements::Int
    recycled_factorizations::Dict{CF,B200LinSolver}
    max_factorizations::Int
end
B200LinSolverCreator(; umfpack_refinements=10, max_factorizations=0) = B200LinSolverCreator(umfpack_refinements, Dict{CF,B200LinSolver}(), max_factorizations)
function create_linsolver(creator::B200LinSolverCreator, nep::B200SPMF, λ)      # LinSolverCreators.jl:107-122
    haskey(creator.recycled_factorizations, CF(λ)) && return creator.recycled_factorizations[CF(λ)]
    solver = B200LinSolver(nep, λ, creator.umfpack_refinements)
    length(creator.recycled_factorizations) < creator.max_factorizations && (creator.recycled_factorizations[CF(λ)] = solver)
    return solver
end
struct B200BackslashLinSolverCreator <: LinSolverCreator end                     # LinSolverCreators.jl:21-37: factorise per solve
struct B200BackslashLinSolver <: LinSolver; nep::B200SPMF; λ::CF; end
create_linsolver(::B200BackslashLinSolverCreator, nep::B200SPMF, λ) = B200BackslashLinSolver(nep, CF(λ))
# `solver.A \ x` on a sparse matrix is `lu(A) \ x` with UMFPACK's default of two refinement steps (LinSolvers.jl:157-159)
lin_solve(s::B200BackslashLinSolver, b::AbstractVecOrMat; tol=0) = lin_solve(B200LinSolver(s.nep, s.λ, 2), b)

# ------------------------------------------------------------------------------------------------
# a9, a10: contour quadrature.  The reference's integrator receives an opaque closure f (method_contour_common.jl:61-94),
# so the batched / sharded path needs the NEP itself: contour_beyn gets a method for B200SPMF that evaluates all N nodes
# in one device call and re-uses the reference's own extraction code through the MatrixIntegrator seam.
# ------------------------------------------------------------------------------------------------
abstract type B200Trapezoidal <: MatrixIntegrator end

"""S[:,:,j] = Σ_i w[i,j] M(λ_i)⁻¹ Vh for this rank's nodes; `reduce=true` sums over the NCCL communicator."""
function b200_contour_moments(nep::B200SPMF, λv::Vector{CF}, W::Matrix{CF}, Vh::Matrix{CF}; batch::Integer=32, reduce::Bool=false,
                              root_only::Bool=false)   # root_only: only rank 0 copies the summed moments back (the others get garbage in S)
    n, k = size(Vh); N, mg = size(W)
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    chk(ccall(sym(:nepb_contour_create), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}), nep.h, k, mg, min(batch, max(N, 1)), ctx))
    coef = Matrix{CF}(undef, length(nep.fi), N)                  # column i = f_1(λ_i)..f_p(λ_i)  (row-major N x p for C)
    for i = 1:N; coef[:, i] = coefficients(nep, λv[i]); end
    Wt = permutedims(W)                                           # mg x N column-major == N x mg row-major
    S = Array{CF,3}(undef, n, k, mg); flags = zeros(Cint, max(N, 1))
    try
        chk(ccall(sym(:nepb_contour_integrate), Cint, (Ptr{Cvoid}, Cint, Ptr{CF}, Ptr{CF}, Ptr{CF}, Int64, Cint, Ptr{CF}, Ptr{Cint}),
                  ctx[], N, coef, Wt, Vh, n, reduce ? (root_only ? 2 : 1) : 0, S, flags))
    finally
        ccall(sym(:nepb_contour_destroy), Cint, (Ptr{Cvoid},), ctx[])
    end
    # bit0 zero pivot, bit1 non-finite pivot, bit2 lifted pivots (after the library's static-pivoting fallback, bit3): the
    # reference would get an accurate UMFPACK solve or a SingularException at such a node, never a silently perturbed one
    any(flags .& 7 .!= 0) && throw(LinearAlgebra.SingularException(0))
    any(flags .& 16 .!= 0) && @warn "quadrature nodes with pivots below 1e-8 max|M_ij|: an eigenvalue lies very close to the contour"
    return S
end

# contour_beyn(nep::B200SPMF; ...) : geometry and weights as in method_beyncontour.jl:69-70,97-111; rank / world shard the nodes
function contour_beyn(::Type{T}, nep::B200SPMF, ::Type{B200Trapezoidal}; σ::Number=zero(complex(T)), radius=1, N::Integer=1000,
                      neigs::Integer=2, k::Integer=neigs + 1, Vh=nothing, batch::Integer=32, rank::Integer=0, world::Integer=1, kwargs...) where {T<:Number}
    n = size(nep, 1)
    r = length(radius) == 1 ? (radius, radius) : radius
    Vhm = Vh === nothing ? Matrix{CF}(randn(n, k)) : Matrix{CF}(Vh)
    h = 2π / N; t = h .* (0:N-1)
    g = complex.(r[1] .* cos.(t), r[2] .* sin.(t)); gp = complex.(-r[1] .* sin.(t), r[2] .* cos.(t))
    W = hcat(gp .* h, gp .* g .* h)                               # temp*G[i,j]*h; the /(2πi) is applied by contour_beyn itself
    mine = (rank+1):world:N
    S = b200_contour_moments(nep, CF.(g[mine] .+ σ), Matrix{CF}(W[mine, :]), Vhm; batch=batch, reduce=world > 1)
    # hand the finished integral to the reference's own extraction code (SVD, rank test, eigen, filters: :110-184).  The
    # integrator is selected by TYPE in the reference (method_contour_common.jl:18-45), so it cannot carry the array itself;
    # task-local storage keeps concurrent contour_beyn calls (one per Task / worker) apart.
    return task_local_storage(:nepb200_precomputed_integral, S) do
        NonlinearEigenproblems.NEPSolver.contour_beyn(T, nep, PrecomputedIntegral; σ=σ, radius=radius, N=N, neigs=neigs, k=k, kwargs...)
    end
end
contour_beyn(nep::B200SPMF; params...) = contour_beyn(CF, nep, B200Trapezoidal; params...)
# integrator that returns an integral computed beforehand (lets the unmodified reference code do the post-processing)
abstract type PrecomputedIntegral <: MatrixIntegrator end
integrate_interval(::Type{PrecomputedIntegral}, ::Type{T}, f, gv, a, b, N, logger) where {T<:Number} =
    task_local_storage(:nepb200_precomputed_integral)::Array{CF,3}

# ------------------------------------------------------------------------------------------------
# (f)3 WEP-native path: new methods for the reference's own WEP_FD type (GalleryWaveguide), src/gallery_extra/waveguide/Waveguide.jl
#   compute_Mlincomb(nep::WEP_FD, λ, V, a)  :324-379   SchurMatVec * v  :393-402   Pinv  :160-163
# The device handle is created once per problem and kept beside it (WeakKeyDict: it dies with the nep).  The reference's
# WEP linear solvers (WEPLinSolverCreator, lin_solve :555-567) run unchanged on top: they only call Pinv, SchurMatVec and a
# factorisation of the Schur complement, for which `B200FactorizeLinSolver` on the assembled matrix is the drop-in.
# ------------------------------------------------------------------------------------------------
# dense n x k block resident in HBM (nepb_block_*): upload / download transpose between Julia's column-major and the device layout
mutable struct B200Block
    h::Ptr{Cvoid}; n::Int; k::Int
    function B200Block(n::Integer, k::Integer)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        chk(ccall(sym(:nepb_block_create), Cint, (Int64, Cint, Ptr{Ptr{Cvoid}}), n, k, h))
        b = new(h[], n, k)
        finalizer(x -> ccall(sym(:nepb_block_destroy), Cint, (Ptr{Cvoid},), x.h), b)
        return b
    end
end
function B200Block(V::Matrix{CF})
    b = B200Block(size(V, 1), size(V, 2))
    chk(ccall(sym(:nepb_block_upload), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{CF}, Int64), b.h, 0, b.k, V, b.n))
    return b
end
function download(b::B200Block)
    V = Matrix{CF}(undef, b.n, b.k)
    chk(ccall(sym(:nepb_block_download), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{CF}, Int64), b.h, 0, b.k, V, b.n))
    return V
end

mutable struct B200WEP
    h::Ptr{Cvoid}
    table_λ::Union{Nothing,CF}; table_cols::Int
end
const WEP_HANDLES = WeakKeyDict{Any,B200WEP}()
function b200_wep(nep)   # nep::GalleryWaveguide.WEP_FD
    get!(WEP_HANDLES, nep) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        K = Matrix{CF}(nep.K); kb = CF[nep.k_bar]; bb = Vector{CF}(nep.bb)
        chk(ccall(sym(:nepb_wep_create), Cint, (Cint, Cint, Cdouble, Cdouble, Ptr{CF}, Ptr{CF}, Ptr{CF}, Ptr{Ptr{Cvoid}}),
                  nep.nx, nep.nz, nep.hx, nep.hz, K, kb, bb, h))
        w = B200WEP(h[], nothing, 0)
        finalizer(x -> ccall(sym(:nepb_wep_destroy), Cint, (Ptr{Cvoid},), x.h), w)
        w
    end
end
# D[m, j] = 1im * d^j/dλ^j sqrt(β_m(λ)) (+ d0 for j = 0) with the reference's own sqrt_derivative (:574-616); row-major for C
function wep_table(nep, λ, ncols::Integer)
    nz = nep.nz; cMP = vcat(nep.cM, nep.cP)
    D = Matrix{CF}(undef, ncols, 2nz)                       # column m of the Julia array = row m of the C array
    for m = 1:2nz
        D[:, m] .= 1im .* GalleryWaveguide.sqrt_derivative(1, nep.b[rem(m - 1, nz) + 1], cMP[m], ncols - 1, λ)
    end
    D[1, :] .+= nep.d0
    return D
end
function b200_wep_mlincomb(nep, λ::Number, V::AbstractVecOrMat, a::Vector=ones(CF, size(V, 2)))
    w = b200_wep(nep); n = size(nep, 1); na = size(V, 2)
    size(V, 1) == n || error("Incompatible sizes: Length of vectors = ", size(V, 1), ", size of NEP = ", n, ".")
    length(a) == na || error("Incompatible sizes: Number of coefficients = ", length(a), ", number of vectors = ", na, ".")
    if w.table_λ != CF(λ) || w.table_cols < na          # a solver loop at a fixed shift uploads the table once
        chk(ccall(sym(:nepb_wep_set_table), Cint, (Ptr{Cvoid}, Cint, Ptr{CF}), w.h, na, wep_table(nep, λ, na)))
        w.table_λ, w.table_cols = CF(λ), na
    end
    Vb = B200Block(Matrix{CF}(reshape(V, n, na))); Zb = B200Block(n, 1)
    chk(ccall(sym(:nepb_wep_mlincomb_block), Cint, (Ptr{Cvoid}, Ptr{CF}, Ptr{Cvoid}, Cint, Cint, Ptr{CF}, Ptr{CF}, Ptr{Cvoid}, Cint),
              w.h, CF[λ], Vb.h, 0, na, Vector{CF}(a), C_NULL, Zb.h, 0))
    return vec(download(Zb))
end
function b200_wep_pinv(nep, λ::Number, x::Vector)
    w = b200_wep(nep); y = similar(x, CF)
    coef = 1 ./ vcat(GalleryWaveguide.sM(nep, λ), GalleryWaveguide.sP(nep, λ))
    chk(ccall(sym(:nepb_wep_pinv), Cint, (Ptr{Cvoid}, Ptr{CF}, Ptr{CF}, Ptr{CF}), w.h, Vector{CF}(coef), Vector{CF}(x), y))
    return y
end
# To activate for the reference's type (needs `using GalleryWaveguide`, an optional sub-package of NEP-PACK):
#   NonlinearEigenproblems.compute_Mlincomb(nep::GalleryWaveguide.WEP_FD, λ::Number, V::AbstractVecOrMat, a::Vector=ones(CF, size(V, 2))) =
#       NEPB200.b200_wep_mlincomb(nep, λ, V, a)
#   GalleryWaveguide.Pinv(nep::GalleryWaveguide.WEP_FD, λ, x) = NEPB200.b200_wep_pinv(nep, λ, x)

# ------------------------------------------------------------------------------------------------
# multi-GPU plumbing: one Julia worker per GPU (`julia -p 8`), NCCL id shipped with Distributed
# ------------------------------------------------------------------------------------------------
function b200_comm_unique_id()
    id = zeros(UInt8, 128); chk(ccall(sym(:nepb_comm_unique_id), Cint, (Ptr{UInt8},), id)); return id
end
function b200_comm_init(nranks::Integer, rank::Integer, id::Vector{UInt8}; device::Integer=rank)
    chk(ccall(sym(:nepb_set_device), Cint, (Cint,), device))
    chk(ccall(sym(:nepb_comm_init), Cint, (Cint, Cint, Ptr{UInt8}), nranks, rank, id))
end

end # module
