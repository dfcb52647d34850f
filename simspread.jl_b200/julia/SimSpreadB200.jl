# SimSpreadB200.jl -- Julia host layer over libsimspread_b200.so (include/simspread_b200.h).
#
# Drop-in for the resource-spreading path of SimSpread.jl: it keeps the exported names and
# signatures of the reference (src/SimSpread.jl:21-56) -- cutoff, cutoff!, featurize, featurize!,
# k, construct, spread, predict, clean!, split, AuROC, AuPRC, recallatL, precisionatL -- and
# forwards every array computation to hand-written sm_100a kernels through `ccall`.  No CUDA.jl
# array dispatch, no CPU fallback.
#
# EXPERIMENTAL: Julia is not installed in the build image nor on the GPU box, so this file has never been executed.
# It is kept mechanical on purpose and mirrors the Python layer `simspread.jl_b200/host.py`, which IS exercised by
# the test-suite through the same C ABI; tests/test_abi_cpu.py checks every ccall against the header (symbol, arity,
# argument widths) and every exported name of the reference's hot path against this file (arity, keyword names).
module SimSpreadB200

using NamedArrays
using Random

export cutoff, cutoff!, featurize, featurize!, k, construct, spread, predict, clean!, save, writedlm,
    AuROC, AuPRC, BEDROC, recallatL, precisionatL, validity_ratio,
    f1score, mcc, accuracy, balancedaccuracy, recall, precision,
    maxperformance, meanperformance, meanstdperformance, jaccard_featurize, recommend_topl, read_namedmatrix,
    predict_blocks, cross_validate, alpha_sweep, Comm, ShardedQuery, predict_sharded

import DelimitedFiles
import DelimitedFiles: writedlm

const libss = get(ENV, "SIMSPREAD_B200_LIB",
    joinpath(@__DIR__, "..", "lib", "libsimspread_b200.so"))

const SS_OP_N = Cint(0)
const SS_OP_T = Cint(1)
const SS_PREDICT_CLEAN = Cuint(1)
const SS_PRECISION = Dict(:f64 => Cuint(0), :tf32 => Cuint(1 << 4), :f64_int8 => Cuint(3 << 4))
const SS_ERR_ASSERT = Cint(4)

# ---------------------------------------------------------------------------------------------
# status handling: reference @assert failures come back as AssertionError with the same text
# ---------------------------------------------------------------------------------------------
function check(status::Cint)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:ss_last_error, libss), Cstring, ()))
    status == SS_ERR_ASSERT ? throw(AssertionError(msg)) : error("libsimspread_b200 status $status: $msg")
end

# ---------------------------------------------------------------------------------------------
# handles (library-owned device memory; finalizers release it)
# ---------------------------------------------------------------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer=parse(Int, get(ENV, "LOCAL_RANK", "0")))
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ss_ctx_create, libss), Cint, (Cint, Ptr{Ptr{Cvoid}}), device, r))
        c = new(r[])
        finalizer(x -> ccall((:ss_ctx_destroy, libss), Cint, (Ptr{Cvoid},), x.h), c)
    end
end
const _ctx = Ref{Union{Nothing,Context}}(nothing)
ctx() = (_ctx[] === nothing && (_ctx[] = Context()); _ctx[]::Context)

mutable struct DMat
    h::Ptr{Cvoid}
    rows::Int
    cols::Int
    function DMat(rows::Integer, cols::Integer)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ss_mat_create, libss), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Ptr{Cvoid}}),
            ctx().h, rows, cols, r))
        m = new(r[], rows, cols)
        finalizer(x -> ccall((:ss_mat_destroy, libss), Cint, (Ptr{Cvoid},), x.h), m)
    end
end

function DMat(A::AbstractMatrix)
    M = Matrix{Float64}(A)                    # column-major, as the ABI expects
    d = DMat(size(M, 1), size(M, 2))
    GC.@preserve M check(ccall((:ss_mat_upload, libss), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64), ctx().h, d.h, M, max(size(M, 1), 1)))
    return d
end

function Base.Matrix(d::DMat)
    M = Matrix{Float64}(undef, d.rows, d.cols)
    GC.@preserve M check(ccall((:ss_mat_download, libss), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64), ctx().h, d.h, M, max(d.rows, 1)))
    return M
end

mutable struct DIVec
    h::Ptr{Cvoid}
    n::Int
    function DIVec(n::Integer)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ss_ivec_create, libss), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}), ctx().h, n, r))
        v = new(r[], n)
        finalizer(x -> ccall((:ss_ivec_destroy, libss), Cint, (Ptr{Cvoid},), x.h), v)
    end
end

function DIVec(idx::AbstractVector{<:Integer}; onebased::Bool=true)
    v32 = Vector{Int32}(onebased ? idx .- 1 : idx)   # ABI indices are 0-based
    d = DIVec(length(v32))
    GC.@preserve v32 check(ccall((:ss_ivec_upload, libss), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}), ctx().h, d.h, v32))
    return d
end

function Base.Vector(d::DIVec)
    v = Vector{Int32}(undef, d.n)
    GC.@preserve v check(ccall((:ss_ivec_download, libss), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}), ctx().h, d.h, v))
    return v
end

hnull(x) = x === nothing ? C_NULL : x.h

# ---------------------------------------------------------------------------------------------
# graphs.jl
# ---------------------------------------------------------------------------------------------
function k(G::AbstractMatrix)
    d = DMat(G)
    out = DIVec(size(G, 1))
    check(ccall((:ss_k_rows, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ctx().h, d.h, out.h))
    return reshape(Int.(Vector(out)), :, 1)             # n x 1 Matrix{Int}, as mapslices(dims=2)
end
k(e::AbstractVector) = k(reshape(e, 1, :))[1]
k(v::Integer, G::AbstractMatrix) = k(G[v, :])

# ---------------------------------------------------------------------------------------------
# core.jl : cutoff / featurize
# ---------------------------------------------------------------------------------------------
function _cutoff(X::AbstractVecOrMat{Float64}, α::Float64, weighted::Bool)
    d = DMat(reshape(X, size(X, 1), :))
    check(ccall((:ss_featurize, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Cint, Ptr{Cvoid}),
        ctx().h, d.h, α, weighted, d.h))
    return reshape(Matrix(d), size(X))
end

cutoff(x::T, α::T, weighted::Bool=false) where {T<:AbstractFloat} =
    T(_cutoff(fill(Float64(x), 1, 1), Float64(α), weighted)[1])
cutoff(X::AbstractVecOrMat{T}, α::T, weighted::Bool=false) where {T<:AbstractFloat} =
    T.(_cutoff(Float64.(X), Float64(α), weighted))
# the reference's cutoff! never mutates its argument (src/core.jl:72-75, 87-89): quirk kept
cutoff!(x::T, α::T, weighted::Bool=false) where {T<:AbstractFloat} = cutoff(x, α, weighted)
cutoff!(X::AbstractVecOrMat{T}, α::T, weighted::Bool=false) where {T<:AbstractFloat} = cutoff(X, α, weighted)

function featurize(X::NamedArray, α::AbstractFloat, weighted::Bool=true)
    X′ = copy(X)
    X′.array = _cutoff(Matrix{Float64}(X.array), Float64(α), weighted)
    setnames!(X′, ["f$f" for f in names(X′, 2)], 2)
    return X′
end

# `featurize(NamedArray(1 .- pairwise(Jaccard(), D, dims=1))[rows, cols], α, weighted)` of the tutorial
# (docs/src/tutorial/fishers-flowers.jl:66, 97-99) in one kernel: D holds one descriptor row per entity and
# the N x N similarity matrix is never materialised (ss_jaccard_featurize).
function jaccard_featurize(D::NamedMatrix, rows::AbstractVector, cols::AbstractVector, α::AbstractFloat,
    weighted::Bool=true)
    da = DMat(Matrix{Float64}(D[rows, :].array))
    db = DMat(Matrix{Float64}(D[cols, :].array))
    X = DMat(length(rows), length(cols))
    check(ccall((:ss_jaccard_featurize, libss), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Cint, Ptr{Cvoid}),
        ctx().h, da.h, db.h, Float64(α), Cint(weighted), X.h))
    return NamedArray(Matrix(X), (string.(rows), ["f$c" for c in cols]))
end

function featurize!(X::NamedArray, α::AbstractFloat, weighted::Bool=true)
    X.array = _cutoff(Matrix{Float64}(X.array), Float64(α), weighted)
    setnames!(X, ["f$f" for f in names(X, 2)], 2)
end

# ---------------------------------------------------------------------------------------------
# core.jl : construct -- index lists instead of the dense n x n matrix
# ---------------------------------------------------------------------------------------------
struct Graph
    queries::Vector{String}
    sources::Vector{String}
    features::Vector{String}
    targets::Vector{String}
    Xq::Union{Nothing,DMat}
    Xs::Union{Nothing,DMat}
    Y::DMat
    masked::Bool      # true for the `B` of the reference (query rows/columns zeroed)
end
Base.names(g::Graph, d::Integer=1) = vcat(g.queries, g.sources, g.features, g.targets)

function _gather(src::DMat, rows, cols, nr, nc)
    dst = DMat(nr, nc)
    check(ccall((:ss_gather, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        ctx().h, src.h, hnull(rows), hnull(cols), dst.h))
    return dst
end

function construct(y::NamedMatrix, X::NamedMatrix, queries::AbstractVector)
    @assert size(y, 1) == size(X, 1) "Labels and features have different number of source nodes"
    features = [f for f in names(X, 2) if lstrip(f, 'f') ∉ queries]
    sources = [d for d in names(X, 1) if d ∉ queries]
    targets = names(y, 2)
    @assert all(sort(features) .!= sort(sources)) "Source and Features nodes have the same names!"
    rowpos = Dict(n => i for (i, n) in enumerate(names(X, 1)))
    colpos = Dict(n => i for (i, n) in enumerate(names(X, 2)))
    yrowpos = Dict(n => i for (i, n) in enumerate(names(y, 1)))
    dX, dy = DMat(X.array), DMat(y.array)
    qi = DIVec([rowpos[q] for q in queries]); si = DIVec([rowpos[s] for s in sources])
    fi = DIVec([colpos[f] for f in features]); ysi = DIVec([yrowpos[s] for s in sources])
    Xq = _gather(dX, qi, fi, length(queries), length(features))
    Xs = _gather(dX, si, fi, length(sources), length(features))
    Y = _gather(dy, ysi, nothing, length(sources), length(targets))
    q, s, f, t = string.(queries), string.(sources), string.(features), string.(targets)
    return Graph(q, s, f, t, Xq, Xs, Y, false), Graph(q, s, f, t, Xq, Xs, Y, true)
end

function construct(ys::T, Xs::T) where {T<:Tuple{NamedMatrix,NamedMatrix}}
    ytrain, ytest = ys
    Xtrain, Xtest = Xs
    @assert size(ytrain, 2) == size(ytest, 2) "Number of targets between test and training sets doesn't match"
    @assert size(Xtrain, 2) == size(Xtest, 2) "Number of features between test and training sets doesn't match"
    features, sources = names(Xtrain, 2), names(ytrain, 1)
    targets, queries = names(ytrain, 2), names(ytest, 1)
    @assert all(sort(features) .!= sort(sources)) "Features and drugs have the same names!"
    Xq, Xsd, Y = DMat(Xtest.array), DMat(Xtrain.array), DMat(ytrain.array)
    q, s, f, t = string.(queries), string.(sources), string.(features), string.(targets)
    return Graph(q, s, f, t, Xq, Xsd, Y, false), Graph(q, s, f, t, Xq, Xsd, Y, true)
end
construct(ytrain::T, ytest::T, Xtrain::T, Xtest::T) where {T<:NamedMatrix} =
    construct((ytrain, ytest), (Xtrain, Xtest))

function construct(y::NamedMatrix, X::NamedMatrix)
    features, sources, targets = names(X, 2), names(y, 1), names(y, 2)
    @assert all(sort(features) .!= sort(sources)) "Source and feature nodes have the same names"
    return Graph(String[], string.(sources), string.(features), string.(targets),
        nothing, DMat(X.array), DMat(y.array), false)
end

# ---------------------------------------------------------------------------------------------
# core.jl : spread / predict / clean!
# ---------------------------------------------------------------------------------------------
function spread(G::AbstractMatrix{Float64})
    d = DMat(G)
    check(ccall((:ss_spread_rows, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        ctx().h, d.h, C_NULL, d.h))
    return Matrix(d)
end
spread(G::AbstractMatrix{Bool}) = spread(Matrix{Float64}(G))
function spread(G::NamedMatrix)
    W = copy(G)
    W.array = spread(Matrix{Float64}(G.array))
    return W
end

# Dense n x n matrix of a Graph, for the forms that are not the block chain (mirrors host.py Graph.array)
function Base.Matrix(g::Graph)
    nq, ns, nf, nt = length(g.queries), length(g.sources), length(g.features), length(g.targets)
    n = nq + ns + nf + nt
    A = zeros(n, n)
    s0, f0, t0 = nq, nq + ns, nq + ns + nf
    if nf > 0
        Xs = Matrix(g.Xs)
        A[s0+1:f0, f0+1:t0] = Xs
        A[f0+1:t0, s0+1:f0] = Xs'
        if nq > 0 && !g.masked
            Xq = Matrix(g.Xq)
            A[1:nq, f0+1:t0] = Xq
            A[f0+1:t0, 1:nq] = Xq'
        end
    end
    Y = Matrix(g.Y)
    A[s0+1:f0, t0+1:n] = Y
    A[t0+1:n, s0+1:f0] = Y'
    return A
end
_named(g::Graph) = NamedArray(Matrix(g), (names(g), names(g)))

# density of the feature blocks from the row-degree kernel (one pass each), not from a CSR build
function _density(M::Union{Nothing,DMat})
    (M === nothing || M.rows * M.cols == 0) && return 0.0
    kk = DIVec(M.rows)
    check(ccall((:ss_k_rows, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ctx().h, M.h, kk.h))
    return sum(Int64.(Vector(kk))) / (M.rows * M.cols)
end

const SPARSE_DENSITY_THRESHOLD = 0.04   # DESIGN.md 4.2: below it the row-split SpMM chain beats the DMMA chain

mutable struct DCsr
    h::Ptr{Cvoid}
    function DCsr(S::DMat, α::Float64, weighted::Bool; by_columns::Bool=false)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        if by_columns
            check(ccall((:ss_featurize_csc, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Cint, Ptr{Ptr{Cvoid}}),
                ctx().h, S.h, α, Cint(weighted), r))
        else
            check(ccall((:ss_featurize_csr, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Cint, Ptr{Ptr{Cvoid}}),
                ctx().h, S.h, α, Cint(weighted), r))
        end
        c = new(r[])
        finalizer(x -> ccall((:ss_csr_destroy, libss), Cint, (Ptr{Cvoid},), x.h), c)
    end
end

# rows `ri` (1-based) and columns `ci` of a device result as a host matrix: one gather kernel, one download
function _fetch(R::DMat, ri::Vector{Int}, ci::Vector{Int})
    (ri == collect(1:R.rows) && ci == collect(1:R.cols)) && return Matrix(R)
    return Matrix(_gather(R, DIVec(ri), DIVec(ci), length(ri), length(ci)))
end

# The block-reduced chain IS the reference's A * (W * W), W = spread(B), only when W has no query edges (B masked, or a
# graph without queries), the rows asked for exist in A (the query rows of a masked A are zero) and A and B are two views
# of one construct() call.  Everything else takes the literal dense path (same rule as host.py predict()).
function _predict_graph(A::Graph, B::Graph, rows, cols; clean::Bool=false, precision::Symbol=:f64, layout::Symbol=:auto)
    layout in (:auto, :dense, :sparse) || throw(ArgumentError("layout must be :auto, :dense or :sparse"))
    haskey(SS_PRECISION, precision) || throw(ArgumentError("precision must be :f64, :f64_int8 or :tf32"))
    flags = (clean ? SS_PREDICT_CLEAN : Cuint(0)) | SS_PRECISION[precision]
    (precision != :f64 && layout == :auto) && (layout = :dense)
    g = A
    qpos = Dict(n => i for (i, n) in enumerate(g.queries))
    spos = Dict(n => i for (i, n) in enumerate(g.sources))
    tpos = Dict(n => i for (i, n) in enumerate(g.targets))
    in_block = all(haskey(tpos, c) for c in cols) && all(haskey(qpos, r) || haskey(spos, r) for r in rows)
    want_q = in_block && any(haskey(qpos, r) for r in rows)
    same_blocks = A.Xs === B.Xs && A.Y === B.Y && A.Xq === B.Xq
    block_ok = in_block && same_blocks && (B.masked || isempty(g.queries)) && !(A.masked && want_q)
    if !block_ok
        return predict((_named(A), _named(B)), NamedArray(zeros(length(rows), length(cols)), (string.(rows), string.(cols))))
    end
    nt = length(g.targets)
    ci = [tpos[c] for c in cols]
    isq = [haskey(qpos, r) for r in rows]
    ri = [isq[i] ? qpos[r] : spos[r] for (i, r) in enumerate(rows)]
    out = Matrix{Float64}(undef, length(rows), length(cols))
    if any(isq)
        R = DMat(length(g.queries), nt)
        if !isempty(g.features) && !isempty(g.sources)
            sparse = layout == :sparse || (layout == :auto && max(_density(g.Xq), _density(g.Xs)) < SPARSE_DENSITY_THRESHOLD)
            if !sparse && all(isq) && ri == collect(1:R.rows) && ci == collect(1:nt)
                # the whole query block on the dense chain: finished column blocks are downloaded while the next ones
                # are computed (out is column-major with leading dimension R.rows)
                GC.@preserve out check(ccall((:ss_predict_query_fetch, libss), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cuint, Ptr{Float64}, Int64),
                    ctx().h, g.Xq.h, g.Xs.h, g.Y.h, R.h, flags, out, Int64(R.rows)))
                return NamedArray(out, (string.(rows), string.(cols)))
            end
            if sparse
                cq = DCsr(g.Xq, -Inf, true)
                cs = DCsr(g.Xs, -Inf, true; by_columns=true)
                check(ccall((:ss_predict_query_csr, libss), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cuint, Ptr{Cvoid}),
                    ctx().h, cq.h, cs.h, g.Y.h, R.h, flags, C_NULL))
            else
                check(ccall((:ss_predict_query, libss), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cuint, Ptr{Cvoid}),
                    ctx().h, g.Xq.h, g.Xs.h, g.Y.h, R.h, flags, C_NULL))
            end
        elseif clean   # no feature layer: the query rows are zero (ss_mat_create zero-fills); clean! still applies
            kt = DIVec(nt)
            check(ccall((:ss_degrees, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                ctx().h, C_NULL, g.Y.h, C_NULL, C_NULL, kt.h))
            check(ccall((:ss_clean, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ctx().h, R.h, kt.h))
        end
        out[isq, :] = _fetch(R, ri[isq], ci)
    end
    if !all(isq)
        R = DMat(length(g.sources), nt)
        check(ccall((:ss_predict_source, libss), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cuint),
            ctx().h, isempty(g.features) ? C_NULL : hnull(g.Xs), g.Y.h, R.h, flags))
        out[.!isq, :] = _fetch(R, ri[.!isq], ci)
    end
    return NamedArray(out, (string.(rows), string.(cols)))
end

# `GPU` is accepted for signature compatibility; the computation always runs on the GPU in Float64
predict(I::Tuple{Graph,Graph}, ytest::NamedMatrix; GPU::Bool=false, clean::Bool=false, precision::Symbol=:f64,
    layout::Symbol=:auto) =
    _predict_graph(I[1], I[2], names(ytest, 1), names(ytest, 2); clean=clean, precision=precision, layout=layout)
predict(A::Graph, B::Graph, ytest::NamedMatrix; GPU::Bool=false, clean::Bool=false, precision::Symbol=:f64,
    layout::Symbol=:auto) = predict((A, B), ytest; GPU=GPU, clean=clean, precision=precision, layout=layout)
predict(A::Graph, ytrain::NamedMatrix; GPU::Bool=false, clean::Bool=false, precision::Symbol=:f64, layout::Symbol=:auto) =
    _predict_graph(A, A, names(ytrain, 1), names(ytrain, 2); clean=clean, precision=precision, layout=layout)

# One call with HOST matrices (what a caller with plain Julia arrays uses; the library pipelines the query slabs
# H2D / GEMM / D2H on three streams): R = Xq * T, T = (Xs' * (Y ./ ks)) ./ kf, clean! fused.
function predict_blocks(Xq::Matrix{Float64}, Xs::Matrix{Float64}, Y::Matrix{Float64}; clean::Bool=true, precision::Symbol=:f64)
    nq, nf = size(Xq)
    ns, nt = size(Y)
    @assert size(Xs) == (ns, nf) "Labels and features have different number of source nodes"
    R = Matrix{Float64}(undef, nq, nt)
    flags = (clean ? SS_PREDICT_CLEAN : Cuint(0)) | SS_PRECISION[precision]
    GC.@preserve Xq Xs Y R check(ccall((:ss_predict_query_host, libss), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int64, Int64, Int64, Cuint, Ptr{Float64}, Int64),
        ctx().h, Xq, max(nq, 1), Xs, max(ns, 1), Y, max(ns, 1), nq, ns, nf, nt, flags, R, max(nq, 1)))
    return R
end

# literal path for arbitrary dense (A, B) NamedArrays: F = A * (W * W), W = spread(B), on the GPU
function predict(I::Tuple{T,T}, ytest::T; GPU::Bool=false) where {T<:NamedMatrix}
    A, B = I
    n = size(A, 1)
    dA, dW = DMat(A.array), DMat(B.array)
    check(ccall((:ss_spread_rows, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        ctx().h, dW.h, C_NULL, dW.h))
    W2, F = DMat(n, n), DMat(n, n)
    gemm(a, b, c) = check(ccall((:ss_gemm_f64, libss), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        ctx().h, SS_OP_N, a.h, b.h, c.h, C_NULL, C_NULL))
    gemm(dW, dW, W2)
    gemm(dA, W2, F)
    Fn = NamedArray(Matrix(F), (names(A, 1), names(A, 2)))
    return Fn[names(ytest, 1), names(ytest, 2)]
end
predict(A::T, B::T, ytest::T; GPU::Bool=false) where {T<:NamedMatrix} = predict((A, B), ytest; GPU=GPU)
predict(A::T, ytrain::T; GPU::Bool=false) where {T<:NamedMatrix} = predict((A, A), ytrain; GPU=GPU)

function clean!(yhat::NamedArray, A::Graph, y::NamedArray)
    kt = DIVec(length(A.targets))
    check(ccall((:ss_degrees, libss), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        ctx().h, C_NULL, A.Y.h, C_NULL, C_NULL, kt.h))
    tpos = Dict(n => i for (i, n) in enumerate(A.targets))
    ktv = Vector(kt)
    for t in names(y, 2)
        ktv[tpos[t]] == 0 && (yhat[:, t] .= -99)
    end
end

function clean!(yhat::NamedArray, A::NamedArray, y::NamedArray)
    for (tᵢ, kᵢ) in zip(names(y, 2), k(A[names(y, 2), :].array))
        kᵢ == 0 && (yhat[:, tᵢ] .= -99)
    end
end

# split is host-only in the reference as well: copied from src/core.jl:11-25 (MIT) so that the folds of a given seed
# are the reference's (the shuffle is Julia's MersenneTwister stream)
function Base.split(y::NamedArray, k::Int64; seed::Int64=1)
    sources = names(y, 1)
    shuffle!(MersenneTwister(seed), sources)
    groups = [[] for _ in 1:k]
    for (i, sᵢ) in enumerate(sources)
        push!(groups[mod(i, k)+1], sᵢ)
    end
    return groups
end

# ---------------------------------------------------------------------------------------------
# performance.jl : ranking metrics
# ---------------------------------------------------------------------------------------------
function _auc(y::AbstractVector{Bool}, yhat::AbstractVector)
    @assert length(y) == length(yhat) "The number of scores must be equal to the number of labels"
    dy, ds = DMat(reshape(Float64.(y), :, 1)), DMat(reshape(Float64.(yhat), :, 1))
    out = zeros(Float64, 2)
    GC.@preserve out check(ccall((:ss_auroc_auprc_mat, libss), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), ctx().h, dy.h, ds.h, out))
    return out
end
AuROC(y::AbstractVector{Bool}, yhat::AbstractVector) = _auc(y, yhat)[1]
AuPRC(y::AbstractVector{Bool}, yhat::AbstractVector) = _auc(y, yhat)[2]

function _atl(Y::AbstractMatrix, S::AbstractMatrix, L::Integer)   # rows = groups
    dy, ds = DMat(Float64.(Y)), DMat(Float64.(S))
    out = zeros(Float64, 2)
    GC.@preserve out check(ccall((:ss_atl, libss), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}), ctx().h, dy.h, ds.h, L, out))
    return out
end

function _atl_grouped(which::Int, y, yhat, grouping, L::Integer)
    @assert length(yhat) == length(grouping) "Number of groups must match number of predictions"
    @assert length(y) == length(grouping) "Number of groups must match number of labels"
    @assert length(y) == length(yhat) "Number of predictions must match number of labels"
    @assert L > 0 "Please use a list length greater than 0 (L > 0)"
    groups = unique(grouping)
    masks = [grouping .== g for g in groups]
    if length(unique(count.(masks))) == 1
        Y = permutedims(reduce(hcat, [y[m] for m in masks]))
        S = permutedims(reduce(hcat, [yhat[m] for m in masks]))
        return _atl(Y, S, L)[which]
    end
    vals = [_atl(reshape(y[m], 1, :), reshape(yhat[m], 1, :), L)[which] for m in masks]
    return sum(vals) / length(vals)
end

recallatL(y, yhat, L::Integer=20) = _atl(reshape(y, 1, :), reshape(yhat, 1, :), L)[1]
precisionatL(y, yhat, L::Integer=20) = _atl(reshape(y, 1, :), reshape(yhat, 1, :), L)[2]
recallatL(y, yhat, grouping, L::Integer=20) = _atl_grouped(1, y, yhat, grouping, L)
precisionatL(y, yhat, grouping, L::Integer=20) = _atl_grouped(2, y, yhat, grouping, L)

validity_ratio(yhat::AbstractVector) = k(yhat) / length(yhat)

function BEDROC(y::AbstractVector{Bool}, yhat::AbstractVector; rev::Bool=true, α::AbstractFloat=20.0)
    @assert length(y) == length(yhat) "The number of scores must be equal to the number of labels"
    dy, ds = DMat(reshape(Float64.(y .== 1), :, 1)), DMat(reshape(Float64.(yhat), :, 1))
    out = Ref{Float64}(0.0)
    check(ccall((:ss_bedroc, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Float64, Ptr{Float64}),
        ctx().h, dy.h, ds.h, rev, Float64(α), out))
    return out[]
end

# metric is one of this module's f1score / mcc / accuracy / balancedaccuracy / recall / precision,
# passed by name: ids follow include/simspread_b200.h
const _METRIC_ID = Dict(:f1score => 0, :mcc => 1, :accuracy => 2, :balancedaccuracy => 3, :recall => 4, :precision => 5)
function _sweep(y::AbstractVector, yhat::AbstractVector, metric::Function)
    @assert length(y) == length(yhat) "The number of scores must be equal to the number of labels"
    dy, ds = DMat(reshape(Float64.(y .!= 0), :, 1)), DMat(reshape(Float64.(yhat), :, 1))
    out = zeros(Float64, 4)
    GC.@preserve out check(ccall((:ss_threshold_sweep, libss), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}), ctx().h, dy.h, ds.h, _METRIC_ID[nameof(metric)], out))
    return out
end
maxperformance(y::AbstractVector, yhat::AbstractVector, metric::Function) = _sweep(y, yhat, metric)[1]
meanperformance(y::AbstractVector, yhat::AbstractVector, metric::Function) = _sweep(y, yhat, metric)[2]
meanstdperformance(y::AbstractVector, yhat::AbstractVector, metric::Function) = (o = _sweep(y, yhat, metric); (o[2], o[3]))

# ---------------------------------------------------------------------------------------------
# sparse 2-layer NBI with fused top-L (BASELINE config 5): `predict(construct(y, X), y)` of the reference on a
# graph without feature layer (src/core.jl:446-466), reduced to `sortperm(rev=true)[1:L]` per source
# (src/performance.jl:315) without materialising the score matrix.  Returns (idx, val): L x sources, 1-based
# target indices (0 = padding) and scores.  `srange` restricts the sources (sharding over GPUs).
# ---------------------------------------------------------------------------------------------
function recommend_topl(y::AbstractMatrix; L::Integer=20, srange::UnitRange{Int}=1:size(y, 1))
    ns, nt = size(y)
    weighted = any(v -> v != 0 && v != 1, y)
    d = DMat(Matrix{Float64}(y))
    α = weighted ? -Inf : 5e-324          # keep every (positive) entry
    hy, hyt = Ref{Ptr{Cvoid}}(C_NULL), Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ss_featurize_csr, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Cint, Ptr{Ptr{Cvoid}}),
        ctx().h, d.h, α, Cint(weighted), hy))
    check(ccall((:ss_featurize_csc, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Cint, Ptr{Ptr{Cvoid}}),
        ctx().h, d.h, α, Cint(weighted), hyt))
    idx, val = DIVec(L * ns), DMat(L, ns)
    try
        check(ccall((:ss_recommend_topl, libss), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
            ctx().h, hy[], hyt[], Cint(L), first(srange) - 1, last(srange), idx.h, val.h))
    finally
        ccall((:ss_csr_destroy, libss), Cint, (Ptr{Cvoid},), hy[])
        ccall((:ss_csr_destroy, libss), Cint, (Ptr{Cvoid},), hyt[])
    end
    return reshape(Vector(idx), L, ns) .+ Int32(1), Matrix(val)
end

# ---------------------------------------------------------------------------------------------
# utils.jl: read_namedmatrix with the value block parsed by all host cores (ss_text_matrix_read); names, the
# "R#i" / "C#j" defaults and the sort by name are the reference's (src/utils.jl:24-53).
# ---------------------------------------------------------------------------------------------
function read_namedmatrix(filepath::String, delimiter::Char=' ', valuetype::Type=Float64; rows::Bool=true, cols::Bool=true)
    nl, nf = Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:ss_text_matrix_dims, libss), Cint, (Cstring, Cint, Ptr{Int64}, Ptr{Int64}),
        filepath, Cint(delimiter), nl, nf))
    nr, nc = nl[] - Int(cols), nf[] - Int(rows)
    vals = Matrix{Float64}(undef, nr, nc)
    GC.@preserve vals check(ccall((:ss_text_matrix_read, libss), Cint,
        (Cstring, Cint, Cint, Cint, Ptr{Float64}, Int64, Int64, Int64),
        filepath, Cint(delimiter), Cint(cols), Cint(rows), vals, nr, nc, max(nr, 1)))
    row_names = ["R#$i" for i in 1:nr]
    col_names = ["C#$j" for j in 1:nc]
    if rows || cols
        for (i, line) in enumerate(eachline(filepath))
            if i == 1 && cols
                col_names = String.(split(line, delimiter))[(rows ? 2 : 1):end]
            elseif rows && i - Int(cols) <= nr
                row_names[i - Int(cols)] = String(first(split(line, delimiter; limit=2)))
            end
            !rows && break
        end
    end
    M = NamedArray(valuetype === Float64 ? vals : valuetype.(vals), (row_names, col_names))
    return M[sort(row_names), sort(col_names)]
end


# ---------------------------------------------------------------------------------------------
# performance.jl:102-296 : scalars of one confusion matrix.  O(1) host arithmetic in the reference too; restated here
# so that `maxperformance(y, yhat, f1score)` works with this module alone (the threshold sweeps themselves run on the
# device, `ss_threshold_sweep`, and pick the metric by the NAME of the function passed).
# ---------------------------------------------------------------------------------------------
_nonempty(tn, fp, fn, tp) = @assert tn + fp + fn + tp > 0 "Confusion matrix sums zero!"
function f1score(tn::T, fp::T, fn::T, tp::T) where {T<:Integer}          # src/performance.jl:102-112
    _nonempty(tn, fp, fn, tp)
    den = tp + 0.5 * (fp + fn)
    return den == 0 ? NaN : tp / den
end
mcc(a::T, b::T, ϵ::AbstractFloat=floatmin(Float64)) where {T<:Integer} =    # :150-152, limit form for an empty row / column
    (a * ϵ - b * ϵ) / sqrt((a + b) * (a + ϵ) * (b + ϵ) * (ϵ + ϵ))
function mcc(tn::T, fp::T, fn::T, tp::T) where {T<:Integer}              # :170-198, branch order as in the reference
    _nonempty(tn, fp, fn, tp)
    p_pred, n_pred, p_act, n_act = tp + fp, fn + tn, tp + fn, fp + tn
    p_pred == 0 && return mcc(tn, fn)
    n_pred == 0 && return mcc(tp, fp)
    p_act == 0 && return mcc(tn, fp)
    n_act == 0 && return mcc(tp, fn)
    return ((tp * tn) - (fp * fn)) / sqrt(p_pred * n_pred * p_act * n_act)
end
function accuracy(tn::T, fp::T, fn::T, tp::T) where {T<:Integer}         # :213-223
    _nonempty(tn, fp, fn, tp)
    den = (tp + tn) + (fp + fn)
    return den == 0 ? NaN : (tp + tn) / den
end
function balancedaccuracy(tn::T, fp::T, fn::T, tp::T) where {T<:Integer} # :240-246
    _nonempty(tn, fp, fn, tp)
    return (tp / (tp + fn) + tn / (tn + fp)) / 2
end
function recall(tn::T, fp::T, fn::T, tp::T) where {T<:Integer}           # :261-270
    _nonempty(tn, fp, fn, tp)
    return tp + fn == 0 ? NaN : tp / (tp + fn)
end
function precision(tn::T, fp::T, fn::T, tp::T) where {T<:Integer}        # :285-294
    _nonempty(tn, fp, fn, tp)
    return tp + fp == 0 ? NaN : tp / (tp + fp)
end

# ---------------------------------------------------------------------------------------------
# core.jl:503-561 `save`, utils.jl:8-11 `writedlm`: rows `fold, "source", "target", score, label` APPENDED to the file.
# The library formats the numbers as Julia's string(x) and writes with all host cores (ss_save_rows).
# ---------------------------------------------------------------------------------------------
function _save(filepath::String, fold::Int64, yhat::NamedMatrix, y::NamedMatrix, delimiter)
    queries, targets = names(y, 1), names(y, 2)
    yh = Matrix{Float64}(yhat[queries, targets].array)
    yy = Matrix{Float64}(y.array)
    qn, tn = string.(queries), string.(targets)
    nbytes = Ref{Int64}(0)
    GC.@preserve yh yy qn tn check(ccall((:ss_save_rows, libss), Cint,
        (Cstring, Cint, Int64, Int64, Int64, Ptr{Cstring}, Ptr{Cstring}, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Int64, Cint, Cint, Ptr{Int64}),
        filepath, Cint(1), fold, length(qn), length(tn), qn, tn, yh, max(size(yh, 1), 1), Cint(eltype(yhat.array) <: Integer),
        yy, max(size(yy, 1), 1), Cint(eltype(y.array) <: Integer), Cint(Char(delimiter)), nbytes))
    return nothing
end
save(filepath::String, yhat::NamedMatrix, y::NamedMatrix; delimiter::Char='\t') = _save(filepath, Int64(-1), yhat, y, delimiter)
save(filepath::String, fidx::Int64, yhat::NamedMatrix, y::NamedMatrix; delimiter='\t') = _save(filepath, fidx, yhat, y, delimiter)

function _namedmatrix2matrix(x::NamedMatrix)     # utils.jl:13-22: names in the first row / column
    M = Matrix{Any}(undef, size(x, 1) + 1, size(x, 2) + 1)
    M[1, 1] = ""
    M[1, 2:end] = names(x, 2)
    M[2:end, 1] = names(x, 1)
    M[2:end, 2:end] = x.array
    return M
end
writedlm(io::IO, x::NamedMatrix{T} where {T}) = writedlm(io, _namedmatrix2matrix(x))
writedlm(io::AbstractString, x::NamedMatrix{T} where {T}) = writedlm(io, _namedmatrix2matrix(x))
writedlm(io::IO, x::NamedMatrix{T} where {T}, delimiter::Char) = writedlm(io, _namedmatrix2matrix(x), delimiter)
writedlm(io::AbstractString, x::NamedMatrix{T} where {T}, delimiter::Char) = writedlm(io, _namedmatrix2matrix(x), delimiter)

# ---------------------------------------------------------------------------------------------
# the loops the reference leaves to its users (docs/src/api.md:17-21), data resident on the GPU
# ---------------------------------------------------------------------------------------------
function _fold_indices(DD::NamedMatrix, DT::NamedMatrix, queries)
    qset = Set(string.(queries))
    rowpos = Dict(n => i for (i, n) in enumerate(names(DD, 1)))
    colpos = Dict(n => i for (i, n) in enumerate(names(DD, 2)))
    yrow = Dict(n => i for (i, n) in enumerate(names(DT, 1)))
    sources = [d for d in names(DD, 1) if !(d in qset)]
    features = [c for c in names(DD, 2) if !(lstrip("f" * c, 'f') in qset)]   # names after featurize: "f" * c (src/core.jl:110, 152)
    qi = Int32[rowpos[q] - 1 for q in queries]
    si = Int32[rowpos[s] - 1 for s in sources]
    fi = Int32[colpos[f] - 1 for f in features]
    ysi = Int32[yrow[s] - 1 for s in sources]
    return qi, si, fi, ysi
end

# k-fold de-novo cross-validation: split -> featurize -> every fold's construct / predict / clean! in ONE library call
# (ss_predict_query_folds: the fold is a grid dimension) -> AuROC / AuPRC over all pairs, mean recall@L / precision@L.
function cross_validate(DT::NamedMatrix, DD::NamedMatrix, α::AbstractFloat; weighted::Bool=true, k::Int=10, seed::Int=1,
    L::Int=20, folds=nothing)
    @assert size(DT, 1) == size(DD, 1) "Labels and features have different number of source nodes"
    folds === nothing && (folds = split(DT, k; seed=seed))
    order = [q for f in folds for q in f]
    dX = DMat(Matrix{Float64}(DD.array))
    check(ccall((:ss_featurize, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Cint, Ptr{Cvoid}),
        ctx().h, dX.h, Float64(α), Cint(weighted), dX.h))
    dy = DMat(Matrix{Float64}(DT.array))
    nt = size(DT, 2)
    R = DMat(length(order), nt)
    qp, sp, fp = Int32[0], Int32[0], Int32[0]
    qa, sa, ysa, fa = Int32[], Int32[], Int32[], Int32[]
    for queries in folds
        qi, si, fi, ysi = _fold_indices(DD, DT, queries)
        append!(qa, qi); append!(sa, si); append!(ysa, ysi); append!(fa, fi)
        push!(qp, qp[end] + length(qi)); push!(sp, sp[end] + length(si)); push!(fp, fp[end] + length(fi))
    end
    GC.@preserve qp qa sp sa ysa fp fa check(ccall((:ss_predict_query_folds, libss), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Cvoid}, Cuint),
        ctx().h, dX.h, dy.h, Cint(length(folds)), qp, qa, sp, sa, ysa, fp, fa, R.h, SS_PREDICT_CLEAN))
    yrow = Dict(n => i for (i, n) in enumerate(names(DT, 1)))
    Yall = _gather(dy, DIVec([yrow[q] for q in order]), nothing, length(order), nt)
    auc = zeros(Float64, 2)
    GC.@preserve auc check(ccall((:ss_auroc_auprc_mat, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
        ctx().h, Yall.h, R.h, auc))
    res = Dict{Symbol,Any}(:folds => folds, :AuROC => auc[1], :AuPRC => auc[2])
    if nt > L
        atl = zeros(Float64, 2)
        GC.@preserve atl check(ccall((:ss_atl, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}),
            ctx().h, Yall.h, R.h, Cint(L), atl))
        res[:recallatL], res[:precisionatL] = atl[1], atl[2]
    end
    res[:yhat] = NamedArray(Matrix(R), (string.(order), string.(names(DT, 2))))
    res[:y] = NamedArray(Matrix(Yall), (string.(order), string.(names(DT, 2))))
    return res
end

# SimSpread's alpha sweep (BASELINE config 3): for every cutoff, featurize -> construct -> predict -> clean! -> metrics
# on the same query set; alpha points are independent: rank r of `world` evaluates alphas[r+1:world:end].
function alpha_sweep(DT::NamedMatrix, DD::NamedMatrix, queries::AbstractVector, alphas::AbstractVector; weighted::Bool=true,
    L::Int=20, rank::Int=0, world::Int=1)
    qs = string.(queries)
    sources = [d for d in names(DD, 1) if !(d in Set(qs))]
    feats = [c for c in names(DD, 2) if !(c in Set(qs))]
    Sq, Ss = Matrix{Float64}(DD[qs, feats].array), Matrix{Float64}(DD[sources, feats].array)
    Y, Yq = Matrix{Float64}(DT[sources, :].array), Matrix{Float64}(DT[qs, :].array)
    out = Dict{Symbol,Any}[]
    for α in alphas[rank+1:world:end]
        Xq, Xs = _cutoff(Sq, Float64(α), weighted), _cutoff(Ss, Float64(α), weighted)
        R = predict_blocks(Xq, Xs, Y; clean=true)
        yb, sc = vec(Yq) .> 0, vec(R)
        push!(out, Dict{Symbol,Any}(:alpha => Float64(α), :AuROC => AuROC(yb, sc), :AuPRC => AuPRC(yb, sc),
            :validity => validity_ratio(sc)))
    end
    return out
end

# ---------------------------------------------------------------------------------------------
# multi-GPU (one Julia process per GPU): every exchange step is behind the C ABI (include/simspread_b200.h 3b).
# `Comm(rank, world; path)` joins the library's NCCL communicator through a file rendezvous (no MPI needed);
# `Comm(rank, world, id)` takes the 128-byte unique id from the host's own channel (`unique_id()` on rank 0).
# ---------------------------------------------------------------------------------------------
function unique_id()
    id = zeros(UInt8, 128)
    GC.@preserve id check(ccall((:ss_comm_unique_id, libss), Cint, (Ptr{UInt8},), id))
    return id
end

mutable struct Comm
    h::Ptr{Cvoid}
    rank::Int
    world::Int
    function Comm(rank::Integer, world::Integer, id::Vector{UInt8})
        r = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve id check(ccall((:ss_comm_init, libss), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}, Ptr{Ptr{Cvoid}}),
            ctx().h, Cint(rank), Cint(world), id, r))
        c = new(r[], rank, world)
        finalizer(x -> ccall((:ss_comm_destroy, libss), Cint, (Ptr{Cvoid},), x.h), c)
    end
    function Comm(rank::Integer, world::Integer; path::String, timeout::Real=120.0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ss_comm_init_file, libss), Cint, (Ptr{Cvoid}, Cint, Cint, Cstring, Float64, Ptr{Ptr{Cvoid}}),
            ctx().h, Cint(rank), Cint(world), path, Float64(timeout), r))
        c = new(r[], rank, world)
        finalizer(x -> ccall((:ss_comm_destroy, libss), Cint, (Ptr{Cvoid},), x.h), c)
    end
end
barrier(c::Comm) = check(ccall((:ss_comm_barrier, libss), Cint, (Ptr{Cvoid},), c.h))

mutable struct ShardedQuery
    h::Ptr{Cvoid}
    comm::Comm
    nt_blk::Int
    function ShardedQuery(comm::Comm, ns::Integer, nf::Integer, nt::Integer)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ss_sharded_create, libss), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Ptr{Cvoid}}), comm.h, ns, nf, nt, r))
        blk, fused = Ref{Int64}(0), Ref{Cint}(0)
        check(ccall((:ss_sharded_info, libss), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Cint}), r[], blk, fused))
        return new(r[], comm, blk[])
    end
end
Base.close(p::ShardedQuery) = (p.h != C_NULL && ccall((:ss_sharded_destroy, libss), Cint, (Ptr{Cvoid},), p.h); p.h = C_NULL; nothing)

# predict for the query rows owned by this rank: Xq = this rank's query-row slab (may be empty), Xs replicated,
# Y = the full label block (the rank's target-column block is cut out here).  Returns R[rank's queries, all targets].
function predict_sharded(p::ShardedQuery, Xq::Matrix{Float64}, Xs::Matrix{Float64}, Y::Matrix{Float64}; clean::Bool=true)
    ns, nt = size(Y)
    c0 = p.comm.rank * p.nt_blk
    Yb = zeros(ns, p.nt_blk)
    c1 = min(nt, c0 + p.nt_blk)
    c1 > c0 && (Yb[:, 1:c1-c0] = Y[:, c0+1:c1])
    dXs, dY = DMat(Xs), DMat(Yb)
    flags = clean ? SS_PREDICT_CLEAN : Cuint(0)
    if size(Xq, 1) == 0
        check(ccall((:ss_predict_query_sharded, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cuint),
            p.h, C_NULL, dXs.h, dY.h, C_NULL, flags))
        return zeros(0, nt)
    end
    dXq, R = DMat(Xq), DMat(size(Xq, 1), nt)
    check(ccall((:ss_predict_query_sharded, libss), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cuint),
        p.h, dXq.h, dXs.h, dY.h, R.h, flags))
    return Matrix(R)
end

end # module
