"""CPU oracle for the SimSpread.jl resource-spreading hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference's algorithm (SimSpread.jl, Julia), written
function by function from the reference sources cited below.  It is the *checker* used by
`tests/`, by `__graft_entry__.smoke()` and by the `cpu_baseline` / `--impl reference` legs of
`bench.py`.  Nothing in the product package (`simspread.jl_b200/`) imports it.

Parity status
-------------
Julia is not installed in this container or on the GPU box, so the reference itself cannot be run.
The oracle is pinned by every golden vector the reference's own test-suite holds for this path
(`test/runtests.jl:20-26, 36-81, 93-99, 103-109, 113-118, 120-158, 160-183, 226-243, 246-266,
280-287`, `test/data/save1..4`) -- see `tests/test_oracle_golden.py`.  AuROC / AuPRC / BEDROC have
NO golden in the reference (placeholders at `test/runtests.jl:210-224`): for those three the oracle
is a restatement of MLBase.roc (0.9.1) and Trapz.trapz (2.0.3) semantics and is **parity unpinned**.

All matrices are NumPy float64; names are Python lists of str.  "file:line" citations are relative
to the reference checkout.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------------
# graphs.jl
# --------------------------------------------------------------------------------------------


def k_vec(e: np.ndarray) -> int:
    """`k(e::AbstractVector) = count(!iszero, e)` (src/graphs.jl:10). NaN counts, -0.0 does not."""
    e = np.asarray(e)
    return int(np.count_nonzero(e != 0))  # NaN != 0 is True, -0.0 != 0 is False


def k_node(v: int, G: np.ndarray) -> int:
    """`k(v, G) = count(!iszero, G[v, :])` (src/graphs.jl:9); `v` is 0-based here."""
    return k_vec(np.asarray(G)[v, :])


def k_mat(G: np.ndarray) -> np.ndarray:
    """`k(G) = mapslices(k, G; dims=2)` (src/graphs.jl:11) -> (n, 1) int64 column."""
    G = np.asarray(G)
    nz = G != 0
    return nz.sum(axis=1, dtype=np.int64).reshape(-1, 1)


# --------------------------------------------------------------------------------------------
# core.jl : cutoff / featurize
# --------------------------------------------------------------------------------------------


def cutoff_scalar(x: float, alpha: float, weighted: bool = False) -> float:
    """src/core.jl:37-43 : `x >= alpha ? (weighted ? x : 1.0) : 0.0` (NaN >= alpha is false)."""
    weight = x if weighted else 1.0
    return weight if x >= alpha else 0.0


def cutoff(X: np.ndarray, alpha: float, weighted: bool = False) -> np.ndarray:
    """src/core.jl:55-60 : element-wise cutoff over a vector or matrix; returns a new array."""
    X = np.asarray(X, dtype=np.float64)
    keep = X >= alpha  # NaN -> False
    if weighted:
        return np.where(keep, X, 0.0)
    return np.where(keep, 1.0, 0.0)


def cutoff_inplace(X: np.ndarray, alpha: float, weighted: bool = False) -> np.ndarray:
    """src/core.jl:72-75, 87-89 : `cutoff!` never mutates its argument (scalar rebinding /
    discarded broadcast) -- it only returns the transformed value.  Quirk kept on purpose."""
    return cutoff(X, alpha, weighted)


def featurize(X: np.ndarray, rownames: Sequence[str], colnames: Sequence[str], alpha: float,
              weighted: bool = True):
    """src/core.jl:106-112 : cutoff every entry (default weighted=true) and prefix column names
    with "f".  Returns (array, rownames, colnames)."""
    return cutoff(X, alpha, weighted), list(rownames), ["f" + str(c) for c in colnames]


# --------------------------------------------------------------------------------------------
# core.jl : split
# --------------------------------------------------------------------------------------------


def split_round_robin(shuffled_sources: Sequence[str], k: int) -> List[List[str]]:
    """src/core.jl:17-24 : element i (1-based) of the *already shuffled* list goes to fold
    `mod(i, k) + 1`.  The shuffle itself (`shuffle!(MersenneTwister(seed), ...)`, :16) is Julia's
    RNG stream and is not reproducible outside Julia; the reference's own test is skipped
    (test/runtests.jl:34)."""
    groups: List[List[str]] = [[] for _ in range(k)]
    for i, s in enumerate(shuffled_sources, start=1):
        groups[(i % k)].append(s)  # fold index mod(i,k)+1 (1-based) == i % k (0-based)
    return groups


# --------------------------------------------------------------------------------------------
# core.jl : construct
# --------------------------------------------------------------------------------------------


def _idx(names: Sequence[str], wanted: Sequence[str]) -> List[int]:
    pos = {}
    for i, n in enumerate(names):
        pos.setdefault(n, i)
    return [pos[w] for w in wanted]


def construct_queries(y, ynames, X, Xnames, queries):
    """src/core.jl:148-201 `construct(y, X, queries)`.

    y : (N, Nt) with names (yrows, ycols);  X : (N, Nfeat) with names (xrows, xcols).
    Returns (A, B, names) with node order queries, sources, features, targets (:192-193);
    B = A with query rows and columns zeroed (:196-198)."""
    y = np.asarray(y, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    yrows, ycols = ynames
    xrows, xcols = Xnames
    assert y.shape[0] == X.shape[0], "Labels and features have different number of source nodes"
    queries = [str(q) for q in queries]
    features = [f for f in xcols if f.lstrip("f") not in queries]  # :152 (strips ALL leading 'f')
    sources = [d for d in xrows if d not in queries]  # :153
    targets = list(ycols)  # :154
    # :156  all(sort(features) .!= sort(sources)) -- element-wise, needs equal lengths
    sf, ss = sorted(features), sorted(sources)
    if len(sf) != len(ss):
        raise ValueError("DimensionMismatch: arrays could not be broadcast to a common size")
    assert all(a != b for a, b in zip(sf, ss)), "Source and Features nodes have the same names!"
    qi, si = _idx(xrows, queries), _idx(xrows, sources)
    fi = _idx(xcols, features)
    ysi, ti = _idx(yrows, sources), _idx(ycols, targets)
    Mqf = X[np.ix_(qi, fi)]
    Msf = X[np.ix_(si, fi)]
    Mst = y[np.ix_(ysi, ti)]
    A = _assemble4(Mqf, Msf, Mst)
    names = queries + sources + features + targets
    B = A.copy()
    nq = len(queries)
    # name-indexed zeroing (:197-198): with unique names this is the first nq rows / columns
    qpos = _idx(names, queries)
    B[qpos, :] = 0.0
    B[:, qpos] = 0.0
    return A, B, names


def _assemble4(Mqf, Msf, Mst):
    """The 16-block hvcat of src/core.jl:165-187 / :240-262."""
    nq, nf = Mqf.shape
    ns, nt = Mst.shape
    assert Msf.shape == (ns, nf)
    n = nq + ns + nf + nt
    A = np.zeros((n, n), dtype=np.float64)
    q0, s0, f0, t0 = 0, nq, nq + ns, nq + ns + nf
    A[q0:s0, f0:t0] = Mqf
    A[s0:f0, f0:t0] = Msf
    A[s0:f0, t0:] = Mst
    A[f0:t0, q0:s0] = Mqf.T
    A[f0:t0, s0:f0] = Msf.T
    A[t0:, s0:f0] = Mst.T
    return A


def construct_split(ytrain, ytrain_names, ytest, ytest_names, Xtrain, Xtrain_names, Xtest,
                    Xtest_names):
    """src/core.jl:217-276 `construct((ytrain,ytest),(Xtrain,Xtest))` and forwarder :294-296."""
    ytrain = np.asarray(ytrain, dtype=np.float64)
    ytest = np.asarray(ytest, dtype=np.float64)
    Xtrain = np.asarray(Xtrain, dtype=np.float64)
    Xtest = np.asarray(Xtest, dtype=np.float64)
    assert ytrain.shape[1] == ytest.shape[1], \
        "Number of targets between test and training sets doesn't match"
    assert Xtrain.shape[1] == Xtest.shape[1], \
        "Number of features between test and training sets doesn't match"
    features = list(Xtrain_names[1])
    sources = list(ytrain_names[0])
    targets = list(ytrain_names[1])
    queries = list(ytest_names[0])
    sf, ss = sorted(features), sorted(sources)
    if len(sf) != len(ss):
        raise ValueError("DimensionMismatch: arrays could not be broadcast to a common size")
    assert all(a != b for a, b in zip(sf, ss)), "Features and drugs have the same names!"
    A = _assemble4(Xtest, Xtrain, ytrain)
    names = [str(x) for x in queries + sources + features + targets]
    B = A.copy()
    qpos = _idx(names, [str(q) for q in queries])
    B[qpos, :] = 0.0
    B[:, qpos] = 0.0
    return A, B, names


def construct_3layer(y, ynames, X, Xnames):
    """src/core.jl:308-337 `construct(y, X)` : [0 X Y; X' 0 0; Y' 0 0], order sources, features,
    targets."""
    y = np.asarray(y, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    features = list(Xnames[1])
    sources = list(ynames[0])
    targets = list(ynames[1])
    sf, ss = sorted(features), sorted(sources)
    if len(sf) != len(ss):
        raise ValueError("DimensionMismatch: arrays could not be broadcast to a common size")
    assert all(a != b for a, b in zip(sf, ss)), "Source and feature nodes have the same names"
    ns, nf = X.shape
    nt = y.shape[1]
    n = ns + nf + nt
    A = np.zeros((n, n), dtype=np.float64)
    A[:ns, ns:ns + nf] = X
    A[:ns, ns + nf:] = y
    A[ns:ns + nf, :ns] = X.T
    A[ns + nf:, :ns] = y.T
    return A, [str(x) for x in sources + features + targets]


# --------------------------------------------------------------------------------------------
# core.jl : spread / predict / clean!
# --------------------------------------------------------------------------------------------


def spread(G: np.ndarray) -> np.ndarray:
    """src/core.jl:365-371 : `W = G ./ k(G)`; Inf -> 0; NaN -> 0 (true division, row-wise)."""
    G = np.asarray(G, dtype=np.float64)
    kk = k_mat(G).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        W = G / kk
    W[np.isinf(W) & (W > 0)] = 0.0  # replace!(W, Inf => 0.0) only replaces +Inf (isequal)
    W[np.isnan(W)] = 0.0
    return W


def predict_dense(A: np.ndarray, B: np.ndarray, names: Sequence[str], rows: Sequence[str],
                  cols: Sequence[str], float32: bool = False) -> np.ndarray:
    """src/core.jl:402-423 `predict((A,B), ytest)`: `W = spread(B)`, `F = A * W^2`
    (`A * (W*W)`, :413), slice `F[names(ytest,1), names(ytest,2)]` (:421).
    `float32=True` mimics `GPU=true` (CuArray{Float32}, :404)."""
    W = spread(B)
    A = np.asarray(A, dtype=np.float64)
    if float32:
        A32, W32 = A.astype(np.float32), W.astype(np.float32)
        F = (A32 @ (W32 @ W32)).astype(np.float64)
    else:
        F = A @ (W @ W)
    ri, ci = _idx(names, rows), _idx(names, cols)
    return F[np.ix_(ri, ci)]


def predict_dense_single(A: np.ndarray, names: Sequence[str], rows: Sequence[str],
                         cols: Sequence[str]) -> np.ndarray:
    """src/core.jl:446-466 `predict(A, ytrain)`: same with `W = spread(A)`."""
    return predict_dense(A, A, names, rows, cols)


def clean(yhat: np.ndarray, A: np.ndarray, names: Sequence[str], targets: Sequence[str]) -> None:
    """src/core.jl:478-484 `clean!`: for every target column t whose row in A has degree 0 set
    `yhat[:, t] = -99` (in place)."""
    ti = _idx(names, targets)
    kk = k_mat(np.asarray(A)[ti, :]).ravel()
    for j, kt in enumerate(kk):
        if kt == 0:
            yhat[:, j] = -99.0


# ---- block-reduced forms (derived from the same lines, verified against the
# ---- literal dense form in tests/test_oracle_golden.py) --------------------------------------


def degrees_blocks(Xs: np.ndarray, Y: np.ndarray):
    """Degrees of B's source / feature / target rows (src/graphs.jl:9-11 on the B of
    src/core.jl:196-198): ks = nnz_row(Xs)+nnz_row(Y), kf = nnz_col(Xs), kt = nnz_col(Y)."""
    nzX = Xs != 0
    nzY = Y != 0
    ks = nzX.sum(axis=1, dtype=np.int64) + nzY.sum(axis=1, dtype=np.int64)
    kf = nzX.sum(axis=0, dtype=np.int64)
    kt = nzY.sum(axis=0, dtype=np.int64)
    return ks, kf, kt


def _div_rows(M: np.ndarray, kk: np.ndarray) -> np.ndarray:
    with np.errstate(divide="ignore", invalid="ignore"):
        W = M / kk.astype(np.float64).reshape(-1, 1)
    W[np.isinf(W) & (W > 0)] = 0.0
    W[np.isnan(W)] = 0.0
    return W


def predict_blocks_query(Xq: np.ndarray, Xs: np.ndarray, Y: np.ndarray) -> np.ndarray:
    """F[q,t] = Xq * T,  T = (Xs' ./ kf) * (Y ./ ks)  -- the only non-zero blocks of
    `A*(W*W)` for query rows (src/core.jl:413 with the A/B of :182-198)."""
    ks, kf, _ = degrees_blocks(Xs, Y)
    Wst = _div_rows(Y, ks)
    Wfs = _div_rows(np.ascontiguousarray(Xs.T), kf)
    T = Wfs @ Wst
    return Xq @ T


def predict_blocks_source(Xs: np.ndarray, Y: np.ndarray) -> np.ndarray:
    """F[s,t] = Xs*T + Y*U, U = (Y' ./ kt) * (Y ./ ks)  (src/core.jl:456 on `construct(y,X)`,
    or source rows of :413)."""
    ks, kf, kt = degrees_blocks(Xs, Y)
    Wst = _div_rows(Y, ks)
    Wfs = _div_rows(np.ascontiguousarray(Xs.T), kf)
    Wts = _div_rows(np.ascontiguousarray(Y.T), kt)
    return Xs @ (Wfs @ Wst) + Y @ (Wts @ Wst)


def two_layer_transfer_loops(Y: np.ndarray):
    """Item x item block of `W^2` for the 2-layer graph `[0 Y; Y' 0]` (BASELINE config 5), literal loops in the
    order the reference's association fixes: `Aarr * Warr^2` (src/core.jl:456) squares W = G ./ k (src/core.jl:366)
    first, so U[t',t] = sum over sources s' (ascending) of fl(Y[s',t']/kt[t']) * fl(Y[s',t]/ks[s']), every product
    and every addition rounded on its own.  Small cases only (pure Python).  Returns the dense U."""
    Y = np.asarray(Y, dtype=np.float64)
    ns, nt = Y.shape
    ks = np.count_nonzero(Y != 0, axis=1)
    kt = np.count_nonzero(Y != 0, axis=0)
    U = np.zeros((nt, nt))
    for tp in range(nt):
        for sp in range(ns):
            if Y[sp, tp] == 0:
                continue
            w1 = Y[sp, tp] / float(kt[tp])
            for t in range(nt):
                if Y[sp, t] != 0:
                    U[tp, t] = U[tp, t] + w1 * (Y[sp, t] / float(ks[sp]))
    return U


def two_layer_scores_loops(Y: np.ndarray) -> np.ndarray:
    """F = Y * U in ascending t' order, F[s,t] = sum_{t' asc} Y[s,t'] * U[t',t] (rows of `Aarr * (Warr^2)` that belong
    to the sources; src/core.jl:456, 464).  Small cases only."""
    Y = np.asarray(Y, dtype=np.float64)
    U = two_layer_transfer_loops(Y)
    ns, nt = Y.shape
    F = np.zeros((ns, nt))
    for s in range(ns):
        for tp in range(nt):
            if Y[s, tp] != 0:
                nzc = np.nonzero(U[tp])[0]
                F[s, nzc] = F[s, nzc] + Y[s, tp] * U[tp, nzc]
    return F


def two_layer_scores_sparse(Y, rows=None):
    """The same sums with scipy.sparse (CSR x CSR Gustavson products add in ascending inner index when the rows are
    sorted, multiply and add rounded separately): U = (Y' ./ kt) * (Y ./ ks), F = Y[rows] * U.  `Y`: dense array or
    scipy sparse matrix (sources x targets).  Returns (F as CSR restricted to `rows`, U as CSR)."""
    import scipy.sparse as sp
    Yc = sp.csr_matrix(Y, dtype=np.float64)
    Yc.eliminate_zeros()
    Yc.sort_indices()
    ks = np.diff(Yc.indptr).astype(np.float64)
    Yt = Yc.T.tocsr()
    Yt.sort_indices()
    kt = np.diff(Yt.indptr).astype(np.float64)
    Wst = Yc.copy()
    Wst.data = Wst.data / np.repeat(ks, np.diff(Yc.indptr))   # true division per element (src/core.jl:366)
    Wts = Yt.copy()
    Wts.data = Wts.data / np.repeat(kt, np.diff(Yt.indptr))
    if rows is None:
        U = Wts @ Wst
        return Yc @ U, U
    # only the rows of U that the requested sources reach (the item order, hence the order of additions, is kept)
    A = Yc[np.asarray(rows)]
    need = np.unique(A.indices)
    U = Wts[need] @ Wst
    A_sub = sp.csr_matrix((A.data, np.searchsorted(need, A.indices), A.indptr), shape=(A.shape[0], need.size))
    return A_sub @ U, U


def recommend_topl(Y, L: int, rows=None):
    """Top-L targets per source under `sortperm(rev=true)` (src/performance.jl:315) of the scores above.
    Returns (idx (n, L) int64, val (n, L))."""
    F, _ = two_layer_scores_sparse(Y, rows)
    n, nt = F.shape
    idx = np.zeros((n, L), dtype=np.int64)
    val = np.zeros((n, L))
    for i in range(n):
        row = np.zeros(nt)
        sl = slice(F.indptr[i], F.indptr[i + 1])
        row[F.indices[sl]] = F.data[sl]
        o = sortperm_rev(row)[:L]
        idx[i] = o
        val[i] = row[o]
    return idx, val


def clean_blocks(R: np.ndarray, kt_full: np.ndarray) -> None:
    """`clean!` (src/core.jl:478-484) in block form: kt_full = degree of each target row of A
    (= nnz of the target's column in the source-target block; query rows have no target edges)."""
    R[:, np.asarray(kt_full) == 0] = -99.0


# --------------------------------------------------------------------------------------------
# core.jl : save
# --------------------------------------------------------------------------------------------


def _jl_num(x) -> str:
    """Julia `string()` of Int / Float64 as used by `join(row, delimiter)` (src/core.jl:519)."""
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    r = repr(x)  # shortest round-trip, same digits as Julia's Ryu
    if "e" in r or "E" in r:
        mant, exp = r.lower().split("e")
        if "." not in mant:
            mant += ".0"
        return f"{mant}e{int(exp)}"
    return r


def save_rows(yhat, y, queries, targets, fold=None, delimiter="\t") -> str:
    """src/core.jl:503-522 (fold column = 1-based index of the query, :512) and :542-561 (fold
    column = the given fold id).  Returns the text that the reference appends to the file."""
    out = []
    for qi, q in enumerate(queries):
        for ti, t in enumerate(targets):
            f = (list(queries).index(q) + 1) if fold is None else fold
            row = [str(f), '"' + q + '"', '"' + t + '"', _jl_num(yhat[qi][ti]), _jl_num(y[qi][ti])]
            out.append(delimiter.join(row) + "\n")
    return "".join(out)


# --------------------------------------------------------------------------------------------
# performance.jl  (third-party semantics from SURVEY.md App. D: MLBase 0.9.1, Trapz 2.0.3)
# --------------------------------------------------------------------------------------------


def _isless_key(v: np.ndarray) -> np.ndarray:
    """uint64 key monotone under Julia `isless` for Float64: -Inf < ... < -0.0 < 0.0 < ... < Inf
    < NaN (all NaNs equal)."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    b = v.view(np.uint64).copy()
    neg = (b >> np.uint64(63)) == 1
    b = np.where(neg, ~b, b | np.uint64(1 << 63))
    b = np.where(np.isnan(v), np.uint64(0xFFFFFFFFFFFFFFFF), b)
    return b


def sortperm_rev(yhat: np.ndarray) -> np.ndarray:
    """`sortperm(yhat; rev=true)`: stable, ties keep ascending original index (App. D)."""
    key = _isless_key(np.asarray(yhat, dtype=np.float64))
    # descending by key, ascending by index among equals: stable sort on inverted key
    return np.argsort(~key, kind="stable")


def roc_counts(y: np.ndarray, yhat: np.ndarray):
    """`thresholds = sort(unique(yhat))`; `roc(y, yhat, thresholds)` (src/performance.jl:53-54,
    78-79; MLBase.roc semantics: predicted positive <=> score >= threshold).  Returns
    (thresholds, tp, fp, P, N) with one entry per unique threshold, ascending."""
    y = np.asarray(y).astype(bool).ravel()
    yhat = np.asarray(yhat, dtype=np.float64).ravel()
    key = _isless_key(yhat)
    order = np.argsort(key, kind="stable")
    ks, ys = key[order], y[order]
    P = int(ys.sum())
    N = int(ys.size - P)
    # start of each run of equal keys
    starts = np.flatnonzero(np.concatenate(([True], ks[1:] != ks[:-1])))
    cpos = np.concatenate(([0], np.cumsum(ys, dtype=np.int64)))
    below_pos = cpos[starts]  # positives with score < threshold
    below_all = starts.astype(np.int64)
    tp = P - below_pos
    fp = N - (below_all - below_pos)
    thr = yhat[order][starts]
    return thr, tp.astype(np.int64), fp.astype(np.int64), P, N


def trapz(x: np.ndarray, y: np.ndarray) -> float:
    """Trapz.trapz(x, y) for vectors (v2.0.3): 0 for length <= 1, else the three-term formula of
    App. D (algebraically the ordinary trapezoid rule)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.size
    if n <= 1:
        return 0.0
    r = (x[1] - x[0]) * y[0] + (x[-1] - x[-2]) * y[-1]
    if n > 2:
        r += float(np.sum((x[2:] - x[:-2]) * y[1:-1]))
    return float(r / 2.0)


def AuROC(y, yhat) -> float:
    """src/performance.jl:49-63 : `abs(trapz(fpr, tpr))`, no (0,0) anchor."""
    assert len(y) == len(yhat), "The number of scores must be equal to the number of labels"
    _, tp, fp, P, N = roc_counts(y, yhat)
    with np.errstate(divide="ignore", invalid="ignore"):
        tpr = tp / np.float64(P)
        fpr = fp / np.float64(N)
    return abs(trapz(fpr, tpr))


def AuPRC(y, yhat) -> float:
    """src/performance.jl:74-89 with SimSpread's own recall (:261-272) / precision (:285-296)."""
    assert len(y) == len(yhat), "The number of scores must be equal to the number of labels"
    _, tp, fp, P, N = roc_counts(y, yhat)
    with np.errstate(divide="ignore", invalid="ignore"):
        rec = np.where(P == 0, np.nan, tp / np.float64(P))
        d = (tp + fp).astype(np.float64)
        prec = np.where(d == 0, np.nan, tp / d)
    return abs(trapz(rec, prec))


def BEDROC(y, yhat, rev: bool = True, alpha: float = 20.0) -> float:
    """src/performance.jl:22-38."""
    y = np.asarray(y).ravel()
    yhat = np.asarray(yhat, dtype=np.float64).ravel()
    assert len(y) == len(yhat), "The number of scores must be equal to the number of labels"
    N = len(y)
    n = int(np.sum(y == 1))
    order = sortperm_rev(yhat) if rev else np.argsort(_isless_key(yhat), kind="stable")
    r = np.flatnonzero(y[order] == 1) + 1  # 1-based ranks
    s = float(np.sum(np.exp(-alpha * r / N)))
    Ra = n / N
    rand_sum = Ra * (1 - math.exp(-alpha)) / (math.exp(alpha / N) - 1)
    fac = Ra * math.sinh(alpha / 2) / (math.cosh(alpha / 2) - math.cosh(alpha / 2 - alpha * Ra))
    cte = 1 / (1 - math.exp(alpha * (1 - Ra)))
    return s * fac / rand_sum + cte


def recallatL(y, yhat, L: int = 20) -> float:
    """src/performance.jl:308-328."""
    assert L > 0, "Please use a list length greater than 0 (L > 0)"
    assert len(y) == len(yhat), "Number of predictions and labels don't match"
    assert len(y) > L, "Number of labels is less than length (L > y)"
    order = sortperm_rev(np.asarray(yhat, dtype=np.float64))
    ys = np.asarray(y)[order]
    Xi = ys.sum()
    XiL = ys[:L].sum()
    return float(XiL / Xi) if Xi > 0 else float("nan")


def precisionatL(y, yhat, L: int = 20) -> float:
    """src/performance.jl:370-385."""
    assert L > 0, "Please use a list length greater than 0 (L > 0)"
    assert len(y) == len(yhat), "Number of predictions and labels don't match"
    assert len(y) > L, "Number of labels is less than length (L > y)"
    order = sortperm_rev(np.asarray(yhat, dtype=np.float64))
    ys = np.asarray(y)[order]
    return float(ys[:L].sum() / L)


def _groups_in_order(grouping):
    seen, out = set(), []
    for g in grouping:
        if g not in seen:
            seen.add(g)
            out.append(g)
    return out


def recallatL_grouped(y, yhat, grouping, L: int = 20) -> float:
    """src/performance.jl:341-357 : mean over groups in order of first appearance; NaN is not
    `missing`, so one all-negative group makes the mean NaN (:356)."""
    y, yhat, grouping = np.asarray(y), np.asarray(yhat, dtype=np.float64), np.asarray(grouping)
    assert len(yhat) == len(grouping), "Number of groups must match number of predictions"
    assert len(y) == len(grouping), "Number of groups must match number of labels"
    assert len(y) == len(yhat), "Number of predictions must match number of labels"
    assert L > 0, "Please use a list length greater than 0 (L > 0)"
    perf = [recallatL(y[grouping == g], yhat[grouping == g], L) for g in _groups_in_order(grouping)]
    return float(np.mean(perf))


def precisionatL_grouped(y, yhat, grouping, L: int = 20) -> float:
    """src/performance.jl:398-414."""
    y, yhat, grouping = np.asarray(y), np.asarray(yhat, dtype=np.float64), np.asarray(grouping)
    assert len(yhat) == len(grouping), "Number of groups must match number of predictions"
    assert len(y) == len(grouping), "Number of groups must match number of labels"
    assert len(y) == len(yhat), "Number of predictions must match number of labels"
    assert L > 0, "Please use a list length greater than 0 (L > 0)"
    perf = [precisionatL(y[grouping == g], yhat[grouping == g], L)
            for g in _groups_in_order(grouping)]
    return float(np.mean(perf))


def validity_ratio(yhat) -> float:
    """src/performance.jl:558-560 : `sum(!iszero, yhat)/length(yhat)` (-99 flags count as valid)."""
    yhat = np.asarray(yhat, dtype=np.float64).ravel()
    return float(np.count_nonzero(yhat != 0) / yhat.size)


# ---- confusion-matrix scalars (src/performance.jl:102-296) -----------------------------------

FLOATMIN = 2.2250738585072014e-308  # floatmin(Float64)


def roc_int(y, pred):
    """`MLBase.roc(gt, pred::IntegerVector)` as used in test/runtests.jl:257-266: pred > 0 is
    positive.  Returns (tn, fp, fn, tp)."""
    y = np.asarray(y) > 0
    p = np.asarray(pred) > 0
    return (int(np.sum(~y & ~p)), int(np.sum(~y & p)), int(np.sum(y & ~p)), int(np.sum(y & p)))


def f1score(tn, fp, fn, tp):
    assert tn + fp + fn + tp > 0, "Confusion matrix sums zero!"
    den = tp + 0.5 * (fp + fn)
    return float("nan") if den == 0 else tp / den


def mcc_eps(a, b, eps=FLOATMIN):
    """src/performance.jl:150-152."""
    return (a * eps - b * eps) / math.sqrt((a + b) * (a + eps) * (b + eps) * (eps + eps))


def mcc(tn, fp, fn, tp):
    """src/performance.jl:170-200 (branch order p_pred, n_pred, p_actual, n_actual)."""
    assert tn + fp + fn + tp > 0, "Confusion matrix sums zero!"
    p_pred, n_pred, p_act, n_act = tp + fp, fn + tn, tp + fn, fp + tn
    if p_pred == 0:
        return mcc_eps(tn, fn)
    if n_pred == 0:
        return mcc_eps(tp, fp)
    if p_act == 0:
        return mcc_eps(tn, fp)
    if n_act == 0:
        return mcc_eps(tp, fn)
    return ((tp * tn) - (fp * fn)) / math.sqrt(p_pred * n_pred * p_act * n_act)


def accuracy(tn, fp, fn, tp):
    assert tn + fp + fn + tp > 0, "Confusion matrix sums zero!"
    den = (tp + tn) + (fp + fn)
    return float("nan") if den == 0 else (tp + tn) / den


def balancedaccuracy(tn, fp, fn, tp):
    assert tn + fp + fn + tp > 0, "Confusion matrix sums zero!"
    with np.errstate(divide="ignore", invalid="ignore"):
        tpr = np.float64(tp) / np.float64(tp + fn)
        tnr = np.float64(tn) / np.float64(tn + fp)
    return float((tpr + tnr) / 2)


def recall(tn, fp, fn, tp):
    assert tn + fp + fn + tp > 0, "Confusion matrix sums zero!"
    p = tp + fn
    return float("nan") if p == 0 else tp / p


def precision(tn, fp, fn, tp):
    assert tn + fp + fn + tp > 0, "Confusion matrix sums zero!"
    d = tp + fp
    return float("nan") if d == 0 else tp / d


def _curve_confusions(y, yhat):
    thr, tp, fp, P, N = roc_counts(y, yhat)
    return [(int(N - f), int(f), int(P - t), int(t)) for t, f in zip(tp, fp)]


def maxperformance(y, yhat, metric):
    """src/performance.jl:425-448."""
    return max(metric(*c) for c in _curve_confusions(y, yhat))


def meanperformance(y, yhat, metric):
    """src/performance.jl:459-489."""
    return float(np.mean([metric(*c) for c in _curve_confusions(y, yhat)]))


def meanstdperformance(y, yhat, metric):
    """src/performance.jl:500-531 (`mean_and_std`: corrected sample std)."""
    v = np.array([metric(*c) for c in _curve_confusions(y, yhat)], dtype=np.float64)
    return float(np.mean(v)), float(np.std(v, ddof=1))


# --------------------------------------------------------------------------------------------
# utils.jl
# --------------------------------------------------------------------------------------------


def read_namedmatrix(text: str, delimiter: str = " ", rows: bool = True, cols: bool = True):
    """`read_namedmatrix(filepath, delimiter, Float64; rows, cols)` (src/utils.jl:50-53) on the contents of a
    file: `readdlm(filepath, delimiter, String)` splits every line at EVERY delimiter (a leading delimiter
    yields an empty first cell -- the corner of the header line), `_parse_matrix` (:24-40) takes the value
    block `M[r_idx:end, c_idx:end]`, the names from the first column / row (or "R#i" / "C#j") and re-orders
    rows and columns by sorted name (:38).  Returns (values, row_names, col_names)."""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    M = [ln.rstrip("\r").split(delimiter) for ln in lines]
    c_idx = 1 if rows else 0
    r_idx = 1 if cols else 0
    values = np.array([[float(v) for v in row[c_idx:]] for row in M[r_idx:]], dtype=np.float64)
    if values.size == 0:
        values = values.reshape(len(M) - r_idx, max(0, (len(M[0]) if M else 0) - c_idx))
    row_names = [row[0] for row in M[r_idx:]] if rows else [f"R#{i + 1}" for i in range(values.shape[0])]
    col_names = list(M[0][c_idx:]) if cols else [f"C#{j + 1}" for j in range(values.shape[1])]
    ro = sorted(range(len(row_names)), key=lambda i: row_names[i])
    co = sorted(range(len(col_names)), key=lambda j: col_names[j])
    return values[np.ix_(ro, co)], [row_names[i] for i in ro], [col_names[j] for j in co]


def namedmatrix2matrix(values, row_names, col_names) -> List[List[object]]:
    """`_namedmatrix2matrix` (src/utils.jl:6): `vcat(["" names(M, 2)...], hcat(names(M, 1), M))`."""
    return [[""] + list(col_names)] + [[r] + list(values[i]) for i, r in enumerate(row_names)]


# --------------------------------------------------------------------------------------------
# upstream similarity (not in src/: the tutorial's user-side step, docs/src/tutorial/fishers-flowers.jl:66)
# --------------------------------------------------------------------------------------------


def jaccard_similarity(XA: np.ndarray, XB: np.ndarray) -> np.ndarray:
    """`1 .- pairwise(Jaccard(), X, dims=1)` (docs/src/tutorial/fishers-flowers.jl:66) between the rows of
    XA and XB.  Distances.jl is not in the tree (a docs dependency); its Jaccard distance accumulates, over
    the descriptors in ascending order, a1 += |x+y| - |x-y| (= 2 min) and a2 += |x+y| + |x-y| (= 2 max) and
    returns 1 - a1/a2, a NaN result (0/0) being replaced by 0.  Pinned BIT-EXACTLY by the shipped
    docs/src/tutorial/data/iris.simmat (= this function of iris.features), see tests/test_oracle_golden.py."""
    XA = np.asarray(XA, dtype=np.float64)
    XB = np.asarray(XB, dtype=np.float64)
    a1 = np.zeros((XA.shape[0], XB.shape[0]))
    a2 = np.zeros_like(a1)
    for kk in range(XA.shape[1]):
        m = np.abs(XA[:, None, kk] - XB[None, :, kk])
        p_ = np.abs(XA[:, None, kk] + XB[None, :, kk])
        a1 += p_ - m
        a2 += p_ + m
    with np.errstate(invalid="ignore", divide="ignore"):
        dist = 1.0 - a1 / a2
    dist = np.where(np.isnan(dist), 0.0, dist)
    return 1.0 - dist


def tanimoto_bits(FA: np.ndarray, FB: np.ndarray) -> np.ndarray:
    """Jaccard index of 0/1 descriptors given as bit-packed uint64 rows: |a & b| / |a | b| (0/0 -> 1,
    the distance-0 convention above)."""
    FA = np.asarray(FA, dtype=np.uint64)
    FB = np.asarray(FB, dtype=np.uint64)
    ba = np.unpackbits(FA.view(np.uint8), axis=1).astype(np.int64)
    bb = np.unpackbits(FB.view(np.uint8), axis=1).astype(np.int64)
    both = ba @ bb.T
    union = ba.sum(1)[:, None] + bb.sum(1)[None, :] - both
    with np.errstate(invalid="ignore", divide="ignore"):
        s_ = both.astype(np.float64) / union.astype(np.float64)
    return np.where(union == 0, 1.0, s_)


# --------------------------------------------------------------------------------------------
# Synthetic workloads -- shared by tests and bench so both arms see the same data
# --------------------------------------------------------------------------------------------


def synth_dense(nq: int, ns: int, nf: int, nt: int, seed: int = 20244, y_density: float = 0.05,
                alpha: float = 0.0, weighted: bool = True):
    """C4-shaped synthetic operands: similarity Uniform[0,1) rounded to 6 digits, thresholded at
    alpha; labels Bernoulli(y_density).  Returns column-major (Fortran) float64 Xq, Xs, Y."""
    rng = np.random.default_rng(seed)
    Sq = np.round(rng.random((nq, nf)), 6)
    Ss = np.round(rng.random((ns, nf)), 6)
    Y = (rng.random((ns, nt)) < y_density).astype(np.float64)
    Xq = np.asfortranarray(cutoff(Sq, alpha, weighted))
    Xs = np.asfortranarray(cutoff(Ss, alpha, weighted))
    return Xq, Xs, np.asfortranarray(Y)
