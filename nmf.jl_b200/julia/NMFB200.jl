# NMFB200.jl -- thin Julia front-end of libnmfb200.so (include/nmfb200.h).
#
# Keeps the reference's API for the accelerated path: `nnmf`, `NMFB200.solve!`, `NMFB200.Result{T}`,
# and the option types `MultUpdate{T}`, `GreedyCD{T}`, `ProjectedALS{T}`, `CoordinateDescent{T}`, `ALSPGrad{T}` (same
# keyword names, defaults and validation as NMF.jl src/multupd.jl:9-43, src/greedycd.jl:10-31, src/projals.jl:18-35,
# src/coorddesc.jl:24-46, src/alspgrad.jl:352-373, src/interf.jl:3-101), so existing code can do
#     const NMF = NMFB200
# All numerics live behind the C ABI; this file only validates, ccalls and maps status -> exception.
# NOTE: Julia is not installed in the build image of this repository, so this file is written against
# the header but has not been executed there; the tested binding is nmf.jl_b200/_lib.py (same calls), and
# tests/test_host_api.py checks every `ccall` argument-type tuple in this file against include/nmfb200.h.
#
# Attribution.  The drop-in contract dictates that the keyword constructors of `MultUpdate` / `GreedyCD` (validation
# statements and messages, from NMF.jl src/multupd.jl:17-42 and src/greedycd.jl:18-30), the argument checks of `nnmf`
# (src/interf.jl:15-37), `Result` with its `==`/`hash` (src/common.jl:21-38) and `nmf_checksize` (src/common.jl:5-16)
# reproduce the reference statement for statement; those ~60 lines are derived from NMF.jl, which is licensed under the
# MIT "Expat" License, Copyright (c) 2014: Dahua Lin and contributors (https://github.com/JuliaStats/NMF.jl, LICENSE.md).
# Everything else in this file, and everything behind the C ABI, is original to this repository.
module NMFB200

using LinearAlgebra
using LinearAlgebra: qr!, svd!
using SparseArrays

export nnmf

const libnmfb200 = get(ENV, "NMFB200_LIB", joinpath(@__DIR__, "..", "libnmfb200.so"))

# ---- status codes (include/nmfb200.h) ----------------------------------------------------------------
const OK, EINVAL, EDIM, ECUDA, ENCCL, ENOMEM, ESTATE, ENOTSUP, ENUMERIC = 0:8

struct CResult            # nmfb200_result
    niters::Int64
    converged::Int32
    engine::Int32
    objvalue::Float64
    last_dev::Float64
    solve_ms::Float64
    upload_ms::Float64
    coordinate_updates::Int64
    kernel_launches::Int64
    hot_kernel_ms::Float64
    hot_kernel_launches::Int64
    sub_iterations::Int64
    tolg_final::Float64
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(device::Integer=0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:nmfb200_create, libnmfb200), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint), ref, device, 0)
        st == OK || error("nmfb200_create failed: ", unsafe_string(ccall((:nmfb200_status_string, libnmfb200), Cstring, (Cint,), st)))
        h = new(ref[])
        finalizer(x -> (x.ptr != C_NULL && ccall((:nmfb200_destroy, libnmfb200), Cint, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), h)
        return h
    end
end

function check(h::Handle, st::Integer)
    st == OK && return
    msg = unsafe_string(ccall((:nmfb200_last_error, libnmfb200), Cstring, (Ptr{Cvoid},), h.ptr))
    st == EINVAL && throw(ArgumentError(msg))
    st == EDIM && throw(DimensionMismatch(msg))
    st == ENUMERIC && throw(LinearAlgebra.PosDefException(1))   # ProjectedALS: Gram not positive definite (the reference ignores potrf!'s info)
    error("libnmfb200 [", unsafe_string(ccall((:nmfb200_status_string, libnmfb200), Cstring, (Cint,), st)), "] ", msg)
end

set_option!(h::Handle, key::AbstractString, value) =
    check(h, ccall((:nmfb200_set_option, libnmfb200), Cint, (Ptr{Cvoid}, Cstring, Cstring), h.ptr, key, string(value)))

# ---- Result (src/common.jl:21-38) ----------------------------------------------------------------------
struct Result{T}
    W::Matrix{T}
    H::Matrix{T}
    niters::Int
    converged::Bool
    objvalue::T
    function Result{T}(W::Matrix{T}, H::Matrix{T}, niters::Int, converged::Bool, objv) where T
        size(W, 2) == size(H, 1) || throw(DimensionMismatch("Inner dimensions of W and H mismatch."))
        new{T}(W, H, niters, converged, objv)
    end
end
Base.:(==)(A::Result, B::Result) = A.W == B.W && A.H == B.H && A.niters == B.niters && A.converged == B.converged && A.objvalue == B.objvalue
Base.hash(s::Result, h::UInt) = hash(s.objvalue, hash(s.converged, hash(s.niters, hash(s.H, hash(s.W, h + (0x09c9f08cfcba6de3 % UInt))))))

# ---- option types --------------------------------------------------------------------------------------
mutable struct MultUpdate{T}            # src/multupd.jl:9-43
    obj::Symbol
    maxiter::Int
    verbose::Bool
    tol::T
    update_H::Bool
    lambda_w::T
    lambda_h::T
    function MultUpdate{T}(; obj::Symbol=:mse, maxiter::Integer=100, verbose::Bool=false, tol::Real=cbrt(eps(T)),
                           update_H::Bool=true, lambda_w::Real=zero(T), lambda_h::Real=zero(T),
                           lambda::Union{Real,Nothing}=nothing) where T
        obj == :mse || obj == :div || throw(ArgumentError("Invalid value for obj."))
        maxiter > 1 || throw(ArgumentError("maxiter must be greater than 1."))
        tol > 0 || throw(ArgumentError("tol must be positive."))
        lambda_w >= 0 || throw(ArgumentError("lambda_w must be non-negative."))
        lambda_h >= 0 || throw(ArgumentError("lambda_h must be non-negative."))
        if lambda !== nothing && lambda >= 0
            @warn "lambda is deprecated, use lambda_w and lambda_h instead."
            lambda_w = iszero(lambda_w) ? lambda : lambda_w
            lambda_h = iszero(lambda_h) ? lambda : lambda_h
        end
        if obj == :div
            lambda_w = max(lambda_w, sqrt(eps(T)))
            lambda_h = max(lambda_h, sqrt(eps(T)))
        end
        new{T}(obj, maxiter, verbose, tol, update_H, lambda_w, lambda_h)
    end
end

mutable struct GreedyCD{T}              # src/greedycd.jl:10-31
    maxiter::Int
    verbose::Bool
    tol::T
    update_H::Bool
    lambda_w::T
    lambda_h::T
    function GreedyCD{T}(; maxiter::Integer=100, verbose::Bool=false, tol::Real=cbrt(eps(T)), update_H::Bool=true,
                         lambda_w::Real=zero(T), lambda_h::Real=zero(T)) where T
        maxiter > 1 || throw(ArgumentError("maxiter must be greater than 1."))
        tol > 0 || throw(ArgumentError("tol must be positive."))
        lambda_w >= 0 || throw(ArgumentError("lambda_w must be non-negative."))
        lambda_h >= 0 || throw(ArgumentError("lambda_h must be non-negative."))
        new{T}(maxiter, verbose, tol, update_H, lambda_w, lambda_h)
    end
end

mutable struct ProjectedALS{T}          # src/projals.jl:18-35 (no validation in the reference)
    maxiter::Int
    verbose::Bool
    tol::T
    update_H::Bool
    lambda_w::T
    lambda_h::T
    ProjectedALS{T}(; maxiter::Integer=100, verbose::Bool=false, tol::Real=cbrt(eps(T)), update_H::Bool=true,
                    lambda_w::Real=cbrt(eps(T)), lambda_h::Real=cbrt(eps(T))) where T =
        new{T}(maxiter, verbose, tol, update_H, lambda_w, lambda_h)
end

mutable struct CoordinateDescent{T}     # src/coorddesc.jl:24-46; `seed` replaces Julia's global RNG for shuffle=true
    maxiter::Int
    verbose::Bool
    tol::T
    update_H::Bool
    α::T
    l₁ratio::T
    regularization::Symbol
    shuffle::Bool
    seed::UInt64
    CoordinateDescent{T}(; maxiter::Integer=100, verbose::Bool=false, tol::Real=cbrt(eps(T)), update_H::Bool=true,
                         α::Real=zero(T), regularization=:both, l₁ratio::Real=zero(T), shuffle::Bool=false,
                         seed::Integer=rand(UInt64)) where T =
        new{T}(maxiter, verbose, tol, update_H, α, l₁ratio, regularization, shuffle, seed)
end

mutable struct ALSPGrad{T}              # src/alspgrad.jl:352-373
    maxiter::Int
    maxsubiter::Int
    tol::T
    tolg::T
    update_H::Bool
    verbose::Bool
    ALSPGrad{T}(; maxiter::Integer=100, maxsubiter::Integer=200, tol::Real=cbrt(eps(T)), tolg::Real=eps(T)^(1/4),
                update_H::Bool=true, verbose::Bool=false) where T =
        new{T}(maxiter, maxsubiter, tol, tolg, update_H, verbose)
end

# ---- set_X / solve! ---------------------------------------------------------------------------------------
for (T, sfx) in ((Float32, "f32"), (Float64, "f64"))
    setx = Symbol("nmfb200_set_X_", sfx)
    @eval function set_X!(h::Handle, X::Matrix{$T}; check_nonneg::Bool=false)
        GC.@preserve X check(h, ccall(($(QuoteNode(setx)), libnmfb200), Cint,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Int64, Int64, Cint), h.ptr, X, size(X, 1), size(X, 2), stride(X, 2), check_nonneg))
    end
    setcsc = Symbol("nmfb200_set_X_csc_", sfx)
    # README.md:22 "Sparse NMF": the three arrays of the SparseMatrixCSC cross PCIe and are expanded on the device
    @eval function set_X!(h::Handle, X::SparseMatrixCSC{$T,Int64}; check_nonneg::Bool=false)
        GC.@preserve X check(h, ccall(($(QuoteNode(setcsc)), libnmfb200), Cint,
            (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{$T}, Int64, Int64, Cint, Cint),
            h.ptr, X.colptr, X.rowval, X.nzval, size(X, 1), size(X, 2), 1, check_nonneg))
    end
    cd = Symbol("nmfb200_solve_cd_", sfx)
    @eval function _solve_cd(h::Handle, W::Matrix{$T}, H::Matrix{$T}, a::CoordinateDescent{$T})
        res = Ref{CResult}()
        reg = a.regularization == :both ? 0 : a.regularization == :components ? 1 : a.regularization == :transformation ? 2 : 3
        GC.@preserve W H check(h, ccall(($(QuoteNode(cd)), libnmfb200), Cint,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Ptr{$T}, Int64, Int64, Int64, $T, $T, $T, Cint, Cint, UInt64, Cint, Cint, Cint, Ref{CResult}),
            h.ptr, W, stride(W, 2), H, stride(H, 2), size(W, 2), a.maxiter, a.tol, a.α, a.l₁ratio, reg, a.shuffle, a.seed,
            a.update_H, a.verbose, 0, res))
        r = res[]
        return Result{$T}(W, H, Int(r.niters), r.converged != 0, $T(r.objvalue))
    end
    pg = Symbol("nmfb200_solve_alspgrad_", sfx)
    @eval function _solve_alspgrad(h::Handle, W::Matrix{$T}, H::Matrix{$T}, a::ALSPGrad{$T})
        res = Ref{CResult}()
        GC.@preserve W H check(h, ccall(($(QuoteNode(pg)), libnmfb200), Cint,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Ptr{$T}, Int64, Int64, Int64, Int64, $T, $T, Cint, Cint, Cint, Ref{CResult}),
            h.ptr, W, stride(W, 2), H, stride(H, 2), size(W, 2), a.maxiter, a.maxsubiter, a.tol, a.tolg, a.update_H, a.verbose, 0, res))
        r = res[]
        return Result{$T}(W, H, Int(r.niters), r.converged != 0, $T(r.objvalue))
    end
    for (alg, name) in ((:multmse, "multmse"), (:multdiv, "multdiv"), (:greedycd, "greedycd"), (:projals, "projals"))
        cname = Symbol("nmfb200_solve_", name, "_", sfx)
        fname = Symbol("_solve_", name)
        @eval function $fname(h::Handle, W::Matrix{$T}, H::Matrix{$T}, maxiter, tol, lw, lh, update_H, verbose)
            res = Ref{CResult}()
            GC.@preserve W H check(h, ccall(($(QuoteNode(cname)), libnmfb200), Cint,
                (Ptr{Cvoid}, Ptr{$T}, Int64, Ptr{$T}, Int64, Int64, Int64, $T, $T, $T, Cint, Cint, Cint, Ref{CResult}),
                h.ptr, W, stride(W, 2), H, stride(H, 2), size(W, 2), maxiter, tol, lw, lh, update_H, verbose, 0, res))
            r = res[]
            return Result{$T}(W, H, Int(r.niters), r.converged != 0, $T(r.objvalue))   # aliases the caller's W, H like the reference
        end
    end
end

# randinit on the device (nmfb200_randinit_*): counter-based Philox keyed by `seed`, for the X resident on the handle
for (T, sfx) in ((Float32, "f32"), (Float64, "f64"))
    rinit = Symbol("nmfb200_randinit_", sfx)
    @eval function randinit!(h::Handle, W::Matrix{$T}, H::Matrix{$T}; seed::Integer=0, normalize::Bool=false, zeroh::Bool=false,
                             row_offset::Integer=0, p_total::Integer=size(W, 1))
        GC.@preserve W H check(h, ccall(($(QuoteNode(rinit)), libnmfb200), Cint,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Ptr{$T}, Int64, Int64, UInt64, Int64, Int64, Cint, Cint, Cint),
            h.ptr, W, stride(W, 2), H, stride(H, 2), size(W, 2), seed % UInt64, row_offset, p_total, normalize, zeroh, 0))
        return W, H
    end
end

function nmf_checksize(X, W::AbstractMatrix, H::AbstractMatrix)   # src/common.jl:5-16
    p, n = size(X); k = size(W, 2)
    (size(W, 1) == p && size(H) == (k, n)) || throw(DimensionMismatch("Dimensions of X, W, and H are inconsistent."))
    return (p, n, k)
end

"""    solve!(alg, X, W, H; handle=Handle()) -> Result{T}   (NMF.solve!, src/multupd.jl:45, src/greedycd.jl:33)"""
function solve!(alg::MultUpdate{T}, X::Union{Matrix{T},SparseMatrixCSC{T,Int64}}, W::Matrix{T}, H::Matrix{T}; handle::Handle=Handle(), x_resident::Bool=false) where T
    nmf_checksize(X, W, H)
    x_resident || set_X!(handle, X)
    f = alg.obj == :mse ? _solve_multmse : _solve_multdiv
    f(handle, W, H, alg.maxiter, alg.tol, alg.lambda_w, alg.lambda_h, alg.update_H, alg.verbose)
end
function solve!(alg::GreedyCD{T}, X::Union{Matrix{T},SparseMatrixCSC{T,Int64}}, W::Matrix{T}, H::Matrix{T}; handle::Handle=Handle(), x_resident::Bool=false) where T
    nmf_checksize(X, W, H)
    x_resident || set_X!(handle, X)
    _solve_greedycd(handle, W, H, alg.maxiter, alg.tol, alg.lambda_w, alg.lambda_h, alg.update_H, alg.verbose)
end

function solve!(alg::ProjectedALS{T}, X::Union{Matrix{T},SparseMatrixCSC{T,Int64}}, W::Matrix{T}, H::Matrix{T}; handle::Handle=Handle(), x_resident::Bool=false) where T
    nmf_checksize(X, W, H)                                    # src/projals.jl:37-39
    x_resident || set_X!(handle, X)
    _solve_projals(handle, W, H, alg.maxiter, alg.tol, alg.lambda_w, alg.lambda_h, alg.update_H, alg.verbose)
end
function solve!(alg::CoordinateDescent{T}, X::Union{Matrix{T},SparseMatrixCSC{T,Int64}}, W::Matrix{T}, H::Matrix{T}; handle::Handle=Handle(), x_resident::Bool=false) where T
    nmf_checksize(X, W, H)                                    # src/coorddesc.jl:49-51
    x_resident || set_X!(handle, X)
    _solve_cd(handle, W, H, alg)
end
function solve!(alg::ALSPGrad{T}, X::Union{Matrix{T},SparseMatrixCSC{T,Int64}}, W::Matrix{T}, H::Matrix{T}; handle::Handle=Handle(), x_resident::Bool=false) where T
    nmf_checksize(X, W, H)                                    # src/alspgrad.jl:381-383
    x_resident || set_X!(handle, X)
    _solve_alspgrad(handle, W, H, alg)
end

"""    solve_batched!(alg::MultUpdate{Float32}, Ws, Hs; handle) -> Vector{Result{Float32}} or `nothing`

`length(Ws)` independent `MultUpdate(obj=:mse)` solves of the X resident on `handle` as ONE stacked iteration
(`nmfb200_solve_multmse_batched_f32`: every pass over X serves all replicates; stop_condition per replicate).  `Ws[r]`, `Hs[r]` are
updated in place.  Returns `nothing` when the library does not cover the request (status ENOTSUP): the caller loops over `solve!`."""
function solve_batched!(alg::MultUpdate{Float32}, Ws::Vector{Matrix{Float32}}, Hs::Vector{Matrix{Float32}}; handle::Handle)
    R = length(Ws); k = size(Ws[1], 2)
    Wst = reduce(hcat, Ws); Hst = reduce(vcat, Hs)            # replicate r: columns / rows (r-1)*k+1 : r*k
    res = Vector{CResult}(undef, R)
    st = GC.@preserve Wst Hst res ccall((:nmfb200_solve_multmse_batched_f32, libnmfb200), Cint,
        (Ptr{Cvoid}, Ptr{Float32}, Int64, Ptr{Float32}, Int64, Int64, Int32, Int64, Float32, Float32, Float32, Cint, Cint, Ref{CResult}),
        handle.ptr, Wst, stride(Wst, 2), Hst, stride(Hst, 2), k, R, alg.maxiter, alg.tol, alg.lambda_w, alg.lambda_h, alg.update_H, 0, res)
    st == ENOTSUP && return nothing
    check(handle, st)
    out = Vector{Result{Float32}}(undef, R)
    for r in 1:R
        copyto!(Ws[r], view(Wst, :, (r-1)*k+1:r*k)); copyto!(Hs[r], view(Hst, (r-1)*k+1:r*k, :))
        out[r] = Result{Float32}(Ws[r], Hs[r], Int(res[r].niters), res[r].converged != 0, Float32(res[r].objvalue))
    end
    return out
end

# ---- randinit (src/initialization.jl:4-17, src/utils.jl:26-32) and nnmf (src/interf.jl:3-101) -------------------
function randinit(p::Integer, n::Integer, k::Integer, T::DataType; normalize::Bool=false, zeroh::Bool=false)
    W = rand(T, p, k)
    if normalize
        for j in 1:k
            W[:, j] .*= 1 / sum(view(W, :, j))
        end
    end
    H = zeroh ? zeros(T, k, n) : rand(T, k, n)
    return W, H
end

# ---- NNDSVD (src/initialization.jl:26-137).  rsvd: RandomizedLinAlg.rsvd(X, k) as called at :78, with its two
# X-sized products on the GPU (nmfb200_mul_X_*, X resident); thin QR and the k x n SVD stay LAPACK calls.
for (T, sfx) in ((Float32, "f32"), (Float64, "f64"))
    fname = "nmfb200_mul_X_" * sfx
    @eval function mul_X(h::Handle, B::Matrix{$T}, rowsC::Integer; transpose::Bool=false)   # rowsC = transpose ? n : p
        C = Matrix{$T}(undef, rowsC, size(B, 2))
        GC.@preserve B C check(h, ccall(($fname, libnmfb200), Cint,
            (Ptr{Cvoid}, Cint, Ptr{$T}, Int64, Int64, Ptr{$T}, Int64),
            h.ptr, transpose ? 1 : 0, B, stride(B, 2), size(B, 2), C, stride(C, 2)))
        return C
    end
end

# NNDSVD entirely on the device (nmfb200_nndsvd_*, csrc/init_device.cuh): Philox Gaussian test matrix keyed by `seed`, CholeskyQR2 and a
# one-sided Jacobi SVD in Float64, `_nndsvd!` as one CTA per component.  Returns `nothing` when the sample is numerically rank deficient
# (status ENUMERIC, k > rank(X)) or the handle is row-sharded (ENOTSUP): the caller then uses the host range finder below.
for (T, sfx) in ((Float32, "f32"), (Float64, "f64"))
    dname = "nmfb200_nndsvd_" * sfx
    @eval function nndsvd_device!(h::Handle, W::Matrix{$T}, H::Matrix{$T}; variant::Symbol=:std, zeroh::Bool=false, seed::Integer=0)
        ivar = variant == :std ? 0 : variant == :a ? 1 : variant == :ar ? 2 : throw(ArgumentError("Invalid value for variant"))
        st = GC.@preserve W H ccall(($dname, libnmfb200), Cint,
            (Ptr{Cvoid}, Ptr{$T}, Int64, Ptr{$T}, Int64, Int64, Cint, Cint, UInt64, Cint),
            h.ptr, W, stride(W, 2), H, stride(H, 2), size(W, 2), ivar, zeroh, seed % UInt64, 0)
        (st == ENUMERIC || st == ENOTSUP) && return nothing
        check(h, st)
        return W, H
    end
end

function rsvd(h::Handle, ::Type{T}, p::Integer, n::Integer, k::Integer) where T
    Q = Matrix(qr!(mul_X(h, randn(T, n, k), p)).Q)          # Y = X * Omega on the GPU
    Bt = mul_X(h, Q, n; transpose=true)                     # B' = X' * Q on the GPU
    F = svd!(Matrix(Bt'))
    return (Q * F.U)[:, 1:k], F.S[1:k], Matrix(F.Vt[1:k, :]')
end

function posnegnorm(x::AbstractArray{T}) where T             # src/initialization.jl:103-115
    pn = zero(T); nn = zero(T)
    for xi in x
        xi > zero(T) ? (pn += abs2(xi)) : (nn += abs2(xi))
    end
    return sqrt(pn), sqrt(nn)
end

function nndsvd(X::AbstractMatrix{T}, k::Integer; zeroh::Bool=false, variant::Symbol=:std, initdata=nothing,
                handle::Union{Handle,Nothing}=nothing) where T
    p, n = size(X)
    ivar = variant == :std ? 0 : variant == :a ? 1 : variant == :ar ? 2 : throw(ArgumentError("Invalid value for variant"))
    U, s, V = if initdata === nothing
        h = handle === nothing ? (h0 = Handle(); set_X!(h0, Matrix{T}(X)); h0) : handle
        rsvd(h, T, p, n, k)
    else
        (initdata.U[:, 1:k], initdata.S[1:k], initdata.V[:, 1:k])
    end
    U = T.(U); s = T.(s); V = T.(V)
    v0 = ivar == 0 ? zero(T) : ivar == 1 ? convert(T, sum(X) / length(X)) : convert(T, sum(X) / length(X) * 0.01)
    W = Matrix{T}(undef, p, k); Ht = Matrix{T}(undef, n, k)
    for j in 1:k                                             # src/initialization.jl:40-67
        x = view(U, :, j); y = view(V, :, j)
        xp, xn = posnegnorm(x); yp, yn = posnegnorm(y)
        mp = xp * yp; mn = xn * yn
        vj = ivar == 2 ? v0 * rand(T) : v0
        if mp >= mn
            ss = sqrt(s[j] * mp)
            W[:, j] .= ifelse.(x .> 0, x .* (ss / xp), vj)
            zeroh || (Ht[:, j] .= ifelse.(y .> 0, y .* (ss / yp), vj))
        else
            ss = sqrt(s[j] * mn)
            W[:, j] .= ifelse.(x .< 0, .-(x .* (ss / xn)), vj)
            zeroh || (Ht[:, j] .= ifelse.(y .< 0, .-(y .* (ss / yn)), vj))
        end
    end
    return W, (zeroh ? zeros(T, k, n) : Matrix(Ht'))
end

function nnmf(X::AbstractMatrix{T}, k::Integer; init::Symbol=:nndsvdar, initdata=nothing, alg::Symbol=:greedycd,
              maxiter::Integer=100, tol::Real=cbrt(eps(T) / 100), replicates::Integer=1,
              W0::Union{AbstractMatrix{T},Nothing}=nothing, H0::Union{AbstractMatrix{T},Nothing}=nothing,
              update_H::Bool=true, verbose::Bool=false, device::Integer=0, seed::Union{Integer,Nothing}=nothing) where T
    # `seed` (an addition): with an integer the NNDSVD initialiser runs entirely on the GPU (nndsvd_device!, counter-based generator)
    eltype(X) <: Number && all(t -> t >= zero(T), X) || throw(ArgumentError("The elements of X must be non-negative."))
    p, n = size(X)
    k <= min(p, n) || throw(ArgumentError("The value of k should not exceed min(size(X))."))
    replicates >= 1 || throw(ArgumentError("The value of replicates must be positive."))
    if !update_H && init != :custom
        @warn "Only W will be updated."
    end
    if init == :custom
        W0 !== nothing && H0 !== nothing || throw(ArgumentError("To use :custom initialization, set W0 and H0."))
        all(t -> t >= zero(T), W0) || throw(ArgumentError("The elements of W0 must be non-negative."))
        size(W0) == (p, k) || throw(ArgumentError("Invalid size for W0."))
        all(t -> t >= zero(T), H0) || throw(ArgumentError("The elements of H0 must be non-negative."))
        size(H0) == (k, n) || throw(ArgumentError("Invalid size for H0."))
    else
        W0 === nothing && H0 === nothing || @warn "Ignore W0 and H0 except for :custom initialization."
    end
    initH = alg != :projals                                  # src/interf.jl:39
    init in (:random, :custom, :nndsvd, :nndsvda, :nndsvdar) || (init == :spa ? error("init=:spa is not on the accelerated path") :
                                                                 throw(ArgumentError("Invalid value for init.")))
    inst = alg == :multmse ? MultUpdate{T}(obj=:mse, maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H) :
           alg == :multdiv ? MultUpdate{T}(obj=:div, maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H) :
           alg == :greedycd ? GreedyCD{T}(maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H) :
           alg == :projals ? ProjectedALS{T}(maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H) :
           alg == :alspgrad ? ALSPGrad{T}(maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H) :
           alg == :cd ? CoordinateDescent{T}(maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H) :
           alg == :spa ? error("alg=:spa is not on the accelerated path") :
           throw(ArgumentError("Invalid algorithm."))
    h = Handle(device)
    Xm = X isa SparseMatrixCSC{T,Int64} ? X : Matrix{T}(X)   # sparse X: only the stored entries cross PCIe (set_X! expands them on the GPU)
    set_X!(h, Xm)                                      # X stays resident on the GPU across init, solve and replicates
    W, H = init == :random ? randinit(p, n, k, T; normalize=true, zeroh=!initH) :
           init == :custom ? (W0::Matrix{T}, H0::Matrix{T}) :     # aliased and updated in place, like `W = W::Matrix{T}` at src/interf.jl:57-58
           begin
               var = init == :nndsvd ? :std : init == :nndsvda ? :a : :ar
               dev = (seed !== nothing && initdata === nothing) ?
                     nndsvd_device!(h, Matrix{T}(undef, p, k), Matrix{T}(undef, k, n); variant=var, zeroh=!initH, seed=seed) : nothing
               dev === nothing ? nndsvd(Xm, k; zeroh=!initH, variant=var, initdata=initdata, handle=h) : dev
           end
    if replicates > 1 && inst isa MultUpdate{Float32} && inst.obj == :mse && !verbose && 2k <= 256
        # src/interf.jl:85-101 in groups of up to 256 ÷ k replicates per stacked iteration (one pass over X per half-step for the group);
        # solve! of MultUpdate draws no random numbers, so drawing a group's restarts up front is the reference's stream
        ret = nothing; rep = 1
        while rep <= replicates
            g = min(256 ÷ k, 32, replicates - rep + 1)
            facs = [r == 1 ? (W, H) : randinit(p, n, k, T; normalize=true, zeroh=!initH) for r in rep:rep+g-1]
            rs = g >= 2 ? solve_batched!(inst, [f[1] for f in facs], [f[2] for f in facs]; handle=h) : nothing
            rs === nothing && (rs = [solve!(inst, Xm, f[1], f[2]; handle=h, x_resident=true) for f in facs])
            for tmp in rs
                (ret === nothing || ret.objvalue > tmp.objvalue) && (ret = tmp)
            end
            rep += g
        end
        return ret
    end
    ret = solve!(inst, Xm, W, H; handle=h, x_resident=true)
    for _ in 2:replicates                              # src/interf.jl:91-98
        Wr, Hr = randinit(p, n, k, T; normalize=true, zeroh=!initH)
        tmp = solve!(inst, Xm, Wr, Hr; handle=h, x_resident=true)
        if ret.objvalue > tmp.objvalue
            ret = tmp
        end
    end
    return ret
end

end # module
