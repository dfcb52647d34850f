# SimSpreadB200.jl -- Julia host layer over libsimspread_b200.so (include/simspread_b200.h).
#
# Drop-in for the resource-spreading path of SimSpread.jl: it keeps the exported names and
# signatures of the reference (src/SimSpread.jl:21-56) -- cutoff, cutoff!, featurize, featurize!,
# k, construct, spread, predict, clean!, split, AuROC, AuPRC, recallatL, precisionatL -- and
# forwards every array computation to hand-written sm_100a kernels through `ccall`.  No CUDA.jl
# array dispatch, no CPU fallback.
#
# NOTE: Julia is not installed in the build image nor on the GPU box, so this file has never been
# executed there.  It is kept mechanical on purpose and mirrors, line for line, the Python layer
# `simspread.jl_b200/host.py`, which *is* exercised by the test-suite through the same C ABI.
module SimSpreadB200

using NamedArrays
using Random

export cutoff, cutoff!, featurize, featurize!, k, construct, spread, predict, clean!,
    AuROC, AuPRC, BEDROC, recallatL, precisionatL, validity_ratio,
    maxperformance, meanperformance, meanstdperformance, jaccard_featurize, recommend_topl, read_namedmatrix

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

function _predict_rows(g::Graph, rows, cols; clean::Bool=false, precision::Symbol=:f64)
    flags = (clean ? SS_PREDICT_CLEAN : Cuint(0)) | SS_PRECISION[precision]
    qpos = Dict(n => i for (i, n) in enumerate(g.queries))
    spos = Dict(n => i for (i, n) in enumerate(g.sources))
    tpos = Dict(n => i for (i, n) in enumerate(g.targets))
    ci = [tpos[c] for c in cols]
    out = Matrix{Float64}(undef, length(rows), length(cols))
    if any(haskey(qpos, r) for r in rows)
        R = DMat(length(g.queries), length(g.targets))
        check(ccall((:ss_predict_query, libss), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cuint, Ptr{Cvoid}),
            ctx().h, g.Xq.h, g.Xs.h, g.Y.h, R.h, flags, C_NULL))
        Rh = Matrix(R)
        for (i, r) in enumerate(rows)
            haskey(qpos, r) && (out[i, :] = Rh[qpos[r], ci])
        end
    end
    if any(haskey(spos, r) for r in rows)
        R = DMat(length(g.sources), length(g.targets))
        check(ccall((:ss_predict_source, libss), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cuint),
            ctx().h, hnull(g.Xs), g.Y.h, R.h, flags))
        Rh = Matrix(R)
        for (i, r) in enumerate(rows)
            haskey(spos, r) && (out[i, :] = Rh[spos[r], ci])
        end
    end
    return NamedArray(out, (string.(rows), string.(cols)))
end

# `GPU` is accepted for signature compatibility; the computation always runs on the GPU in Float64
function predict(I::Tuple{Graph,Graph}, ytest::NamedMatrix; GPU::Bool=false, clean::Bool=false,
    precision::Symbol=:f64)
    A, _ = I
    return _predict_rows(A, names(ytest, 1), names(ytest, 2); clean=clean, precision=precision)
end
predict(A::Graph, B::Graph, ytest::NamedMatrix; GPU::Bool=false, clean::Bool=false, precision::Symbol=:f64) =
    predict((A, B), ytest; GPU=GPU, clean=clean, precision=precision)
predict(A::Graph, ytrain::NamedMatrix; GPU::Bool=false, clean::Bool=false, precision::Symbol=:f64) =
    _predict_rows(A, names(ytrain, 1), names(ytrain, 2); clean=clean, precision=precision)

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

# split is host-only in the reference as well (src/core.jl:11-25); reproduced verbatim in spirit
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

end # module
