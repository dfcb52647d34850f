"""Host-side mirror of SimSpread.jl's exported API for the resource-spreading path
(reference src/SimSpread.jl:21-56), written in Python because Julia is not available in the build
image; `julia/SimSpreadB200.jl` is the same layer in Julia over the same C ABI.

Same names, argument meaning and error behaviour as the reference: `cutoff`, `cutoff_` (cutoff!),
`featurize`, `featurize_` (featurize!), `k`, `construct`, `spread`, `predict`, `clean_` (clean!),
`split`, `AuROC`, `AuPRC`, `recallatL`, `precisionatL`.  Index arguments that are 1-based in Julia
(`k(v, G)`) are 1-based here too.  All array work is done by libsimspread_b200.so on the GPU; this
module only does name bookkeeping (the NamedArrays part of the reference) and argument checks.
There is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import SS_OP_N, SS_OP_T, SS_PREDICT_CLEAN, check, lib
from .namedarray import NamedArray

# ------------------------------------------------------------------------------------------------
# device plumbing
# ------------------------------------------------------------------------------------------------


class Context:
    """One GPU (`ss_ctx`).  `Context.default()` picks LOCAL_RANK (one process per GPU)."""

    _default: Optional["Context"] = None

    def __init__(self, device: Optional[int] = None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        check(lib().ss_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    @classmethod
    def default(cls) -> "Context":
        if cls._default is None:
            cls._default = Context()
        return cls._default

    def sync(self):
        check(lib().ss_ctx_sync(self.h))

    def stream(self) -> int:
        s = C.c_void_p()
        check(lib().ss_ctx_stream(self.h, C.byref(s)))
        return s.value or 0

    def launch_count(self) -> int:
        n = C.c_int64()
        check(lib().ss_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def int8_stats(self):
        """(products on the INT8 tensor pipe, products re-run in FP64 after a failed certificate, entries that failed
        in the last product) -- bookkeeping of precision="f64_int8"."""
        out = (C.c_int64 * 4)()
        check(lib().ss_ctx_int8_stats(self.h, out))
        return int(out[0]), int(out[1]), int(out[2])

    def int8_last_pairs(self) -> int:
        """Slice pairs of the last int8 product (21 for two general operands with 6 slices, 6 when one is 0/1)."""
        out = (C.c_int64 * 4)()
        check(lib().ss_ctx_int8_stats(self.h, out))
        return int(out[3])

    def profile(self, enable: bool):
        check(lib().ss_ctx_profile(self.h, int(enable)))

    def profile_read(self, cap: int = 4096):
        """[(ms, flops)] of every chain-product GEMM launched since profile(True)."""
        ms, fl, n = (C.c_double * cap)(), (C.c_double * cap)(), C.c_int32()
        check(lib().ss_ctx_profile_read(self.h, ms, fl, cap, C.byref(n)))
        return [(ms[i], fl[i]) for i in range(n.value)]

    def close(self):
        """Destroys the context.  Handles created on it (DMat / DIVec / DCsr) notice (`ctx.h is None`) and skip their
        own destroy call instead of handing the library a dangling context."""
        if self.h:
            lib().ss_ctx_destroy(self.h)
            self.h = None
            if Context._default is self:
                Context._default = None


# below this size the host-side transposing copy is cheaper than a temporary device buffer + one more kernel + a sync
# (measured on the 445 x 664 matrices of C2: 3.7 ms per CV with the host copy, 4.9 ms through the device transpose)
_ROWMAJOR_UPLOAD_MIN_BYTES = 16 << 20


def _f64_colmajor(a) -> np.ndarray:
    a = np.asarray(a)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a, dtype=np.float64)


class DMat:
    """Device matrix handle (`ss_mat`), column-major float64."""

    def __init__(self, ctx: Context, rows: int, cols: int, ipc: bool = False):
        self.ctx = ctx
        self.rows, self.cols = int(rows), int(cols)
        h = C.c_void_p()
        fn = lib().ss_mat_create_ipc if ipc else lib().ss_mat_create
        check(fn(ctx.h, self.rows, self.cols, C.byref(h)))
        self.h = h

    @classmethod
    def from_host(cls, ctx: Context, a) -> "DMat":
        a = np.asarray(a)
        if (a.ndim == 2 and a.dtype == np.float64 and a.flags.c_contiguous and not a.flags.f_contiguous
                and a.shape[0] <= 2_000_000 and a.nbytes >= _ROWMAJOR_UPLOAD_MIN_BYTES):
            # NumPy's default (row-major) order: uploaded as it lies and transposed on the device instead of a strided
            # transposing copy on the host (np.asfortranarray: 0.3 s for the 800 MB similarity matrix of C3)
            m = cls(ctx, a.shape[0], a.shape[1])
            check(lib().ss_mat_upload_rowmajor(ctx.h, m.h, a.ctypes.data, a.shape[1]))
            return m
        a = _f64_colmajor(a)
        m = cls(ctx, a.shape[0], a.shape[1])
        if a.size:
            check(lib().ss_mat_upload(ctx.h, m.h, a.ctypes.data, max(a.shape[0], 1)))
        return m

    @classmethod
    def wrap(cls, ctx: Context, devptr: int, rows: int, cols: int, ld: int) -> "DMat":
        m = cls.__new__(cls)
        m.ctx, m.rows, m.cols = ctx, int(rows), int(cols)
        h = C.c_void_p()
        check(lib().ss_mat_wrap(ctx.h, C.c_void_p(devptr), m.rows, m.cols, int(ld), C.byref(h)))
        m.h = h
        return m

    def info(self):
        r, c, ld, p = C.c_int64(), C.c_int64(), C.c_int64(), C.c_void_p()
        check(lib().ss_mat_info(self.h, C.byref(r), C.byref(c), C.byref(ld), C.byref(p)))
        return r.value, c.value, ld.value, p.value

    def to_host(self) -> np.ndarray:
        out = np.empty((self.rows, self.cols), dtype=np.float64, order="F")
        if out.size:
            check(lib().ss_mat_download(self.ctx.h, self.h, out.ctypes.data, max(self.rows, 1)))
        return out

    def __del__(self):
        try:
            if getattr(self, "h", None) and getattr(self.ctx, "h", None):  # a closed context took its memory with it
                lib().ss_mat_destroy(self.h)
            self.h = None
        except Exception:
            pass


class DIVec:
    """Device int32 vector handle (`ss_ivec`)."""

    def __init__(self, ctx: Context, n: int):
        self.ctx, self.n = ctx, int(n)
        h = C.c_void_p()
        check(lib().ss_ivec_create(ctx.h, self.n, C.byref(h)))
        self.h = h

    @classmethod
    def from_host(cls, ctx: Context, a) -> "DIVec":
        a = np.ascontiguousarray(a, dtype=np.int32)
        v = cls(ctx, a.size)
        if a.size:
            check(lib().ss_ivec_upload(ctx.h, v.h, a.ctypes.data))
        return v

    def to_host(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.int32)
        if self.n:
            check(lib().ss_ivec_download(self.ctx.h, self.h, out.ctypes.data))
        return out

    def __del__(self):
        try:
            if getattr(self, "h", None) and getattr(self.ctx, "h", None):
                lib().ss_ivec_destroy(self.h)
            self.h = None
        except Exception:
            pass


class DCsr:
    """Device CSR handle (`ss_csr`)."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h
        r, c, n, v = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        check(lib().ss_csr_info(h, C.byref(r), C.byref(c), C.byref(n), C.byref(v)))
        self.rows, self.cols, self.nnz, self.has_values = r.value, c.value, n.value, bool(v.value)

    @classmethod
    def from_dense(cls, ctx: Context, S: "DMat", alpha: float, weighted: bool, by_columns: bool = False) -> "DCsr":
        """Threshold + compact (`ss_featurize_csr`), or the CSR of S' with by_columns=True
        (`ss_featurize_csc`)."""
        h = C.c_void_p()
        fn = lib().ss_featurize_csc if by_columns else lib().ss_featurize_csr
        check(fn(ctx.h, S.h, float(alpha), int(bool(weighted)), C.byref(h)))
        return cls(ctx, h)

    @property
    def density(self) -> float:
        return self.nnz / max(1, self.rows * self.cols)

    def to_host(self):
        rp = np.empty(self.rows + 1, np.int32)
        ci = np.empty(max(self.nnz, 1), np.int32)
        va = np.empty(max(self.nnz, 1), np.float64)
        check(lib().ss_csr_download(self.ctx.h, self.h, rp.ctypes.data, ci.ctypes.data,
                                    va.ctypes.data if self.has_values else None))
        return rp, ci[:self.nnz], (va[:self.nnz] if self.has_values else None)

    def __del__(self):
        try:
            if getattr(self, "h", None) and getattr(self.ctx, "h", None):
                lib().ss_csr_destroy(self.h)
            self.h = None
        except Exception:
            pass


# Below this density of both feature blocks the row-split SpMM chain beats the DMMA chain: one
# partial product costs 8 B of L2/HBM traffic (~1.2e-12 s at 6.5 TB/s) against 2 flop of a dense
# FP64 GEMM (~5.5e-14 s at 36 TFLOP/s) -> break-even near 4.5 %.
SPARSE_DENSITY_THRESHOLD = 0.04


def _h(x):
    return x.h if x is not None else None


# ------------------------------------------------------------------------------------------------
# graphs.jl : k
# ------------------------------------------------------------------------------------------------


def k(*args):
    """`k(v, G)` (1-based row), `k(e::Vector)`, `k(G::Matrix)` -- reference src/graphs.jl:9-11.
    `k(G)` returns an (n, 1) int64 column like `mapslices(k, G; dims=2)`."""
    ctx = Context.default()
    if len(args) == 2:
        v, G = args
        G = G.array if isinstance(G, NamedArray) else np.asarray(G)
        return int(k(np.asarray(G)[int(v) - 1, :]))
    (G,) = args
    G = G.array if isinstance(G, NamedArray) else np.asarray(G)
    if G.ndim == 1:  # one row
        d = DMat.from_host(ctx, G.reshape(1, -1))
        out = DIVec(ctx, 1)
        check(lib().ss_k_rows(ctx.h, d.h, out.h))
        return int(out.to_host()[0])
    d = DMat.from_host(ctx, G)
    out = DIVec(ctx, G.shape[0])
    check(lib().ss_k_rows(ctx.h, d.h, out.h))
    return out.to_host().astype(np.int64).reshape(-1, 1)


# ------------------------------------------------------------------------------------------------
# core.jl : cutoff / featurize
# ------------------------------------------------------------------------------------------------


def _cutoff_array(X, alpha: float, weighted: bool) -> np.ndarray:
    ctx = Context.default()
    X = np.asarray(X)
    shape = X.shape
    d = DMat.from_host(ctx, X)
    check(lib().ss_featurize(ctx.h, d.h, float(alpha), int(bool(weighted)), d.h))
    return d.to_host().reshape(shape, order="F") if X.ndim == 1 else d.to_host()


def cutoff(x, alpha, weighted: bool = False):
    """reference src/core.jl:37-43 (scalar) and :55-60 (vector / matrix); default weighted=false.
    The reference requires `typeof(x) == typeof(alpha)` (both AbstractFloat)."""
    if np.isscalar(x):
        if not isinstance(x, (float, np.floating)) or not isinstance(alpha, (float, np.floating)):
            raise TypeError("MethodError: cutoff(x::T, alpha::T) needs AbstractFloat x and alpha")
        return float(_cutoff_array(np.array([[x]], dtype=np.float64), alpha, weighted)[0, 0])
    if not isinstance(alpha, (float, np.floating)):
        raise TypeError("MethodError: cutoff(X::AbstractVecOrMat{T}, alpha::T) needs AbstractFloat alpha")
    return _cutoff_array(x, alpha, weighted)


def cutoff_(x, alpha, weighted: bool = False):
    """`cutoff!` (reference src/core.jl:72-75, 87-89): despite the name the reference never
    mutates its argument -- it only returns the transformed value.  Quirk kept."""
    return cutoff(x, alpha, weighted)


def featurize(X: NamedArray, alpha, weighted: bool = True) -> NamedArray:
    """reference src/core.jl:106-112: threshold every entry (default weighted=true) and prefix the
    column names with "f"."""
    out = X.copy()
    out.array = _cutoff_array(X.array, float(alpha), weighted)
    out.setnames(["f" + f for f in out.names(2)], 2)
    return out


def featurize_(X: NamedArray, alpha, weighted: bool = True) -> None:
    """`featurize!` (reference src/core.jl:129-132): in place."""
    X.array = _cutoff_array(X.array, float(alpha), weighted)
    X.setnames(["f" + f for f in X.names(2)], 2)


# ------------------------------------------------------------------------------------------------
# core.jl : split
# ------------------------------------------------------------------------------------------------


def jaccard_featurize(D: NamedArray, rows: Sequence[str], cols: Sequence[str], alpha, weighted: bool = True) -> NamedArray:
    """`featurize(S[rows, cols], alpha, weighted)` with `S = 1 .- pairwise(Jaccard(), D, dims=1)` (the
    tutorial's similarity step, docs/src/tutorial/fishers-flowers.jl:66, then src/core.jl:106-112) in one
    kernel: `D` holds one descriptor row per entity, the similarity matrix is never materialised."""
    ctx = Context.default()
    ri = D.index_of(rows, 1)
    ci = D.index_of(cols, 1)
    da = DMat.from_host(ctx, D.array[ri, :])
    db = DMat.from_host(ctx, D.array[ci, :])
    X = DMat(ctx, len(ri), len(ci))
    check(lib().ss_jaccard_featurize(ctx.h, da.h, db.h, float(alpha), int(bool(weighted)), X.h))
    return NamedArray(X.to_host(), (list(rows), ["f" + str(c) for c in cols]))


def tanimoto_featurize_bits(FA, FB, alpha, weighted: bool = True) -> np.ndarray:
    """Same for bit-packed fingerprints (uint64 words, one row per entity): X[i,j] = cutoff(|a&b| / |a|b|)."""
    import torch  # device buffers only
    ctx = Context.default()
    FA = np.ascontiguousarray(FA, dtype=np.uint64)
    FB = np.ascontiguousarray(FB, dtype=np.uint64)
    assert FA.ndim == 2 and FB.ndim == 2 and FA.shape[1] == FB.shape[1], "fingerprints must have the same number of words"
    dev = torch.device("cuda", ctx.device)
    ta = torch.from_numpy(FA.view(np.int64)).to(dev)
    tb = torch.from_numpy(FB.view(np.int64)).to(dev)
    X = DMat(ctx, FA.shape[0], FB.shape[0])
    check(lib().ss_tanimoto_featurize_bits(ctx.h, C.c_void_p(ta.data_ptr()), FA.shape[0], C.c_void_p(tb.data_ptr()),
                                           FB.shape[0], FA.shape[1], float(alpha), int(bool(weighted)), X.h))
    return X.to_host()


def split(y: NamedArray, k_: int, seed: int = 1) -> List[List[str]]:
    """reference src/core.jl:11-25: shuffle the source names, element i (1-based) -> fold
    `mod(i, k) + 1`.  The shuffle uses NumPy's MT19937 stream, not Julia's MersenneTwister stream
    (not reproducible outside Julia; the reference's own test is skipped, test/runtests.jl:34)."""
    sources = y.names(1)
    perm = np.random.RandomState(seed).permutation(len(sources))
    groups: List[List[str]] = [[] for _ in range(k_)]
    for i, j in enumerate(perm, start=1):
        groups[i % k_].append(sources[j])  # mod(i,k)+1 in 1-based fold numbering
    return groups


# ------------------------------------------------------------------------------------------------
# core.jl : construct
# ------------------------------------------------------------------------------------------------


class Graph:
    """What `construct` returns in place of the reference's dense n x n NamedArray: the three
    non-zero blocks (Xq = A[q,f], Xs = A[s,f], Y = A[s,t]) resident on the GPU plus the node names
    in the reference's order (queries, sources, features, targets; src/core.jl:192-193).
    `masked=True` is the `B` of the reference (query rows / columns zeroed, :196-198).  It still
    satisfies the `(A, B)` protocol of `predict` and exposes `.names(d)` / `.array` (the dense
    matrix, materialised on request for inspection only)."""

    def __init__(self, ctx, queries, sources, features, targets, Xq: Optional[DMat], Xs: Optional[DMat],
                 Y: DMat, masked: bool):
        self.ctx = ctx
        self.queries, self.sources = list(queries), list(sources)
        self.features, self.targets = list(features), list(targets)
        self.Xq, self.Xs, self.Y = Xq, Xs, Y
        self.masked = masked
        self._csr = None

    def csr(self):
        """(CSR of Xq, CSR of Xs') of the already-featurized blocks, built on first use."""
        if self._csr is None:
            ninf = float("-inf")  # keep every stored non-zero of the featurized block
            self._csr = (DCsr.from_dense(self.ctx, self.Xq, ninf, True),
                         DCsr.from_dense(self.ctx, self.Xs, ninf, True, by_columns=True))
        return self._csr

    def density(self):
        """(density of Xq, density of Xs) from the row-degree kernel (`ss_k_rows`, one pass over each block)."""
        if getattr(self, "_density", None) is None:
            out = []
            for M in (self.Xq, self.Xs):
                if M is None or M.rows * M.cols == 0:
                    out.append(0.0)
                    continue
                kk = DIVec(self.ctx, M.rows)
                check(lib().ss_k_rows(self.ctx.h, M.h, kk.h))
                out.append(float(kk.to_host().sum(dtype=np.int64)) / float(M.rows * M.cols))
            self._density = tuple(out)
        return self._density

    def names(self, d: Optional[int] = None):
        n = [str(x) for x in self.queries + self.sources + self.features + self.targets]
        return [n, list(n)] if d is None else n

    @property
    def array(self) -> np.ndarray:
        nq, ns, nf, nt = len(self.queries), len(self.sources), len(self.features), len(self.targets)
        n = nq + ns + nf + nt
        A = np.zeros((n, n))
        s0, f0, t0 = nq, nq + ns, nq + ns + nf
        if nf:
            Xs = self.Xs.to_host()
            A[s0:f0, f0:t0] = Xs
            A[f0:t0, s0:f0] = Xs.T
            if nq and not self.masked:
                Xq = self.Xq.to_host()
                A[:nq, f0:t0] = Xq
                A[f0:t0, :nq] = Xq.T
        Y = self.Y.to_host()
        A[s0:f0, t0:] = Y
        A[t0:, s0:f0] = Y.T
        return A

    def to_named(self) -> NamedArray:
        return NamedArray(self.array, self.names())


def _assert_names_differ(features, sources, msg):
    # reference: @assert all(sort(features) .!= sort(sources)) msg   (element-wise broadcast)
    sf, ss = sorted(features), sorted(sources)
    if len(sf) != len(ss):
        raise ValueError("DimensionMismatch: arrays could not be broadcast to a common size; "
                         f"got a dimension with lengths {len(sf)} and {len(ss)}")
    assert all(a != b for a, b in zip(sf, ss)), msg


def construct(*args):
    """`construct(y, X, queries)` (reference src/core.jl:148-201), `construct((ytrain, ytest),
    (Xtrain, Xtest))` (:217-276), `construct(ytrain, ytest, Xtrain, Xtest)` (:294-296) and the
    3-layer `construct(y, X)` (:308-337)."""
    if len(args) == 4:
        return construct((args[0], args[1]), (args[2], args[3]))
    if len(args) == 3:
        y, X, queries = args
        assert y.size(1) == X.size(1), "Labels and features have different number of source nodes"
        queries = [str(q) for q in queries]
        qset = set(queries)
        features = [f for f in X.names(2) if f.lstrip("f") not in qset]
        sources = [d for d in X.names(1) if d not in qset]
        targets = y.names(2)
        _assert_names_differ(features, sources, "Source and Features nodes have the same names!")
        ctx = Context.default()
        dX, dy = DMat.from_host(ctx, X.array), DMat.from_host(ctx, y.array)
        qi = DIVec.from_host(ctx, X.index_of(queries, 1))
        si = DIVec.from_host(ctx, X.index_of(sources, 1))
        fi = DIVec.from_host(ctx, X.index_of(features, 2))
        ysi = DIVec.from_host(ctx, y.index_of(sources, 1))
        Xq, Xs = DMat(ctx, len(queries), len(features)), DMat(ctx, len(sources), len(features))
        Y = DMat(ctx, len(sources), len(targets))
        check(lib().ss_gather(ctx.h, dX.h, qi.h, fi.h, Xq.h))
        check(lib().ss_gather(ctx.h, dX.h, si.h, fi.h, Xs.h))
        check(lib().ss_gather(ctx.h, dy.h, ysi.h, None, Y.h))
        A = Graph(ctx, queries, sources, features, targets, Xq, Xs, Y, masked=False)
        B = Graph(ctx, queries, sources, features, targets, Xq, Xs, Y, masked=True)
        return A, B
    if len(args) == 2 and isinstance(args[0], tuple):
        (ytrain, ytest), (Xtrain, Xtest) = args
        assert ytrain.size(2) == ytest.size(2), \
            "Number of targets between test and training sets doesn't match"
        assert Xtrain.size(2) == Xtest.size(2), \
            "Number of features between test and training sets doesn't match"
        features, sources = Xtrain.names(2), ytrain.names(1)
        targets, queries = ytrain.names(2), ytest.names(1)
        _assert_names_differ(features, sources, "Features and drugs have the same names!")
        ctx = Context.default()
        Xq, Xs = DMat.from_host(ctx, Xtest.array), DMat.from_host(ctx, Xtrain.array)
        Y = DMat.from_host(ctx, ytrain.array)
        A = Graph(ctx, queries, sources, features, targets, Xq, Xs, Y, masked=False)
        B = Graph(ctx, queries, sources, features, targets, Xq, Xs, Y, masked=True)
        return A, B
    if len(args) == 2:
        y, X = args
        features, sources, targets = X.names(2), y.names(1), y.names(2)
        _assert_names_differ(features, sources, "Source and feature nodes have the same names")
        ctx = Context.default()
        Xs, Y = DMat.from_host(ctx, X.array), DMat.from_host(ctx, y.array)
        return Graph(ctx, [], sources, features, targets, None, Xs, Y, masked=False)
    raise TypeError("MethodError: no method matching construct(...)")


# ------------------------------------------------------------------------------------------------
# core.jl : spread / predict / clean!
# ------------------------------------------------------------------------------------------------


def spread(G):
    """reference src/core.jl:365-380: `W = G ./ k(G)`, Inf -> 0, NaN -> 0.  Accepts a Float64 /
    Bool matrix or a NamedArray (returns the same kind)."""
    ctx = Context.default()
    if isinstance(G, Graph):
        G = G.to_named()
    arr = G.array if isinstance(G, NamedArray) else np.asarray(G)
    d = DMat.from_host(ctx, arr.astype(np.float64))
    check(lib().ss_spread_rows(ctx.h, d.h, None, d.h))
    W = d.to_host()
    if isinstance(G, NamedArray):
        out = G.copy()
        out.array = W
        return out
    return W


def _dense_predict(A: NamedArray, B: NamedArray, rows, cols) -> NamedArray:
    """Literal reference path for arbitrary dense adjacency matrices: F = A * (W * W) with
    W = spread(B) (src/core.jl:408-413), all three steps on the GPU."""
    ctx = Context.default()
    n = A.shape[0]
    dA, dB = DMat.from_host(ctx, A.array), DMat.from_host(ctx, B.array)
    check(lib().ss_spread_rows(ctx.h, dB.h, None, dB.h))  # dB <- W
    W2, F = DMat(ctx, n, n), DMat(ctx, n, n)
    check(lib().ss_gemm_f64(ctx.h, SS_OP_N, dB.h, dB.h, W2.h, None, None))
    check(lib().ss_gemm_f64(ctx.h, SS_OP_N, dA.h, W2.h, F.h, None, None))
    Fn = NamedArray(F.to_host(), A.names())
    return Fn[list(rows), list(cols)]


_PRECISION_FLAGS = {"f64": _lib.SS_PRECISION_F64, "tf32": _lib.SS_PRECISION_TF32,
                    "f64_int8": _lib.SS_PRECISION_F64_INT8}


def predict(*args, GPU: bool = False, clean: bool = False, layout: str = "auto",
            precision: str = "f64") -> NamedArray:
    """`predict((A, B), ytest)`, `predict(A, B, ytest)` (reference src/core.jl:402-425) and
    `predict(A, ytrain)` (:446-466).  Returns `F[names(ytest,1), names(ytest,2)]`.

    `GPU` is accepted for signature compatibility: this implementation always runs on the GPU, in
    Float64 (the reference's GPU=true silently drops to Float32, src/core.jl:404).  `clean=True`
    (extension) fuses `clean!` into the product's epilogue.  `precision` (extension): "f64" (default,
    DMMA) or "tf32" (tcgen05.mma kind::tf32 with FP32 accumulation in TMEM -- the fast, reduced
    precision analogue of the reference's Float32 GPU mode; relative error ~2e-4).  `layout` (extension) selects the dense
    DMMA chain, the sparse row-split SpMM chain, or picks by the density of the feature blocks."""
    if layout not in ("auto", "dense", "sparse"):
        raise ValueError("layout must be 'auto', 'dense' or 'sparse'")
    if precision not in _PRECISION_FLAGS:
        raise ValueError("precision must be 'f64' (default), 'f64_int8' or 'tf32' (tcgen05, opt-in)")
    ctx = Context.default()
    if len(args) == 2 and isinstance(args[0], tuple):
        (A, B), yq = args
    elif len(args) == 3:
        A, B, yq = args
    elif len(args) == 2:
        A, yq = args
        B = A
    else:
        raise TypeError("MethodError: no method matching predict(...)")
    rows, cols = yq.names(1), yq.names(2)
    if not isinstance(A, Graph) or not isinstance(B, Graph):
        An = A.to_named() if isinstance(A, Graph) else A
        Bn = B.to_named() if isinstance(B, Graph) else B
        return _dense_predict(An, Bn, rows, cols)
    g = A
    flags = (SS_PREDICT_CLEAN if clean else 0) | _PRECISION_FLAGS[precision]
    if precision != "f64" and layout == "auto":
        layout = "dense"  # the reduced-precision modes exist for the dense tensor-core chain only
    qpos = {n: i for i, n in enumerate(g.queries)}
    spos = {n: i for i, n in enumerate(g.sources)}
    tpos = {n: i for i, n in enumerate(g.targets)}
    in_block = all(c in tpos for c in cols) and all((r in qpos) or (r in spos) for r in rows)
    want_q = any(r in qpos for r in rows) if in_block else False
    # The block-reduced chain IS the reference's A * (W * W) only when W = spread(B) has no query
    # edges (B masked, or a graph without queries), when the rows asked for exist in A (query rows of a masked A are
    # zero) and when A and B are two views of the same construct() call.  Everything else -- predict(A, y) on the
    # unmasked 4-layer graph (its feature degrees count the query edges too), predict((B, B), yq), graphs of two
    # different construct() calls, names outside the (query | source) x target block -- takes the literal dense path.
    same_blocks = A.Xs is B.Xs and A.Y is B.Y and A.Xq is B.Xq
    block_ok = in_block and same_blocks and (B.masked or not g.queries) and not (A.masked and want_q)
    if not block_ok:
        return _dense_predict(A.to_named(), B.to_named(), rows, cols)
    nt = len(g.targets)
    ci = np.fromiter((tpos[c] for c in cols), dtype=np.int32, count=len(cols))
    is_q = np.fromiter((r in qpos for r in rows), dtype=bool, count=len(rows))
    ri = np.fromiter((qpos[r] if r in qpos else spos[r] for r in rows), dtype=np.int32, count=len(rows))
    all_cols = len(cols) == nt and np.array_equal(ci, np.arange(nt, dtype=np.int32))

    def fetch(R: DMat, row_sel: np.ndarray) -> np.ndarray:
        """rows `row_sel`, columns `ci` of the device result as a host array: gathered on the device (one kernel), one
        download -- no per-row host loop."""
        if all_cols and len(row_sel) == R.rows and np.array_equal(row_sel, np.arange(R.rows, dtype=np.int32)):
            return R.to_host()
        sub = DMat(ctx, len(row_sel), len(ci))
        dri = DIVec.from_host(ctx, row_sel)            # keep the handles alive across the call
        dci = None if all_cols else DIVec.from_host(ctx, ci)
        check(lib().ss_gather(ctx.h, R.h, dri.h, _h(dci), sub.h))
        return sub.to_host()

    out = None
    if is_q.any():
        R = DMat(ctx, len(g.queries), nt)
        whole = (is_q.all() and all_cols and len(rows) == R.rows
                 and np.array_equal(ri, np.arange(R.rows, dtype=np.int32)))
        if len(g.features) and len(g.sources):
            use_sparse = layout == "sparse"
            if layout == "auto":  # density from the degree kernels (one pass), not from a CSR build
                use_sparse = max(g.density()) < SPARSE_DENSITY_THRESHOLD
            if not use_sparse and whole:
                # the whole query block, dense chain: the download of finished column blocks overlaps the product
                out = np.empty((R.rows, nt), order="F")
                check(lib().ss_predict_query_fetch(ctx.h, g.Xq.h, g.Xs.h, g.Y.h, R.h, flags, out.ctypes.data, R.rows))
                g.last_layout = "dense"
                return NamedArray(out, (rows, cols))
            if use_sparse:
                cq, cs = g.csr()
                check(lib().ss_predict_query_csr(ctx.h, cq.h, cs.h, g.Y.h, R.h, flags, None))
            else:
                check(lib().ss_predict_query(ctx.h, g.Xq.h, g.Xs.h, g.Y.h, R.h, flags, None))
            g.last_layout = "sparse" if use_sparse else "dense"
        else:  # no feature layer / no sources: the query rows of A * (W * W) are zero (ss_mat_create zero-fills)
            if clean:
                kt = DIVec(ctx, nt)
                check(lib().ss_degrees(ctx.h, None, g.Y.h, None, None, kt.h))
                check(lib().ss_clean(ctx.h, R.h, kt.h))
        part = fetch(R, ri[is_q])
        if is_q.all():
            out = part
        else:
            out = np.empty((len(rows), len(cols)), order="F")
            out[is_q, :] = part
    if not is_q.all():
        R = DMat(ctx, len(g.sources), nt)
        check(lib().ss_predict_source(ctx.h, _h(g.Xs) if len(g.features) else None, g.Y.h, R.h, flags))
        part = fetch(R, ri[~is_q])
        if out is None:
            out = part
        else:
            out[~is_q, :] = part
    return NamedArray(out, (rows, cols))


def clean_(yhat: NamedArray, A, y: NamedArray) -> None:
    """`clean!(yhat, A, y)` (reference src/core.jl:478-484): every target column whose node has
    degree 0 in A is flagged with -99, in place."""
    ctx = Context.default()
    targets = y.names(2)
    if isinstance(A, Graph):
        kt = DIVec(ctx, len(A.targets))
        check(lib().ss_degrees(ctx.h, None, A.Y.h, None, None, kt.h))
        pos = {n: i for i, n in enumerate(A.targets)}
        ktv = kt.to_host()[[pos[t] for t in targets]]
    else:
        rows = A[targets, A.names(2)]
        ktv = k(rows.array).ravel()
    dk = DIVec.from_host(ctx, ktv.astype(np.int32))
    sub = yhat[yhat.names(1), targets]
    d = DMat.from_host(ctx, sub.array)
    check(lib().ss_clean(ctx.h, d.h, dk.h))
    res = d.to_host()
    ci = yhat.index_of(targets, 2)
    yhat.array = np.array(yhat.array, dtype=np.float64)
    yhat.array[:, ci] = res


# ------------------------------------------------------------------------------------------------
# performance.jl : ranking metrics
# ------------------------------------------------------------------------------------------------


def _auc_pair(y, yhat) -> Tuple[float, float]:
    ctx = Context.default()
    y = np.asarray(y).ravel()
    yhat = np.asarray(yhat, dtype=np.float64).ravel()
    assert len(y) == len(yhat), "The number of scores must be equal to the number of labels"
    dy = DMat.from_host(ctx, (y != 0).astype(np.float64).reshape(-1, 1))
    ds = DMat.from_host(ctx, yhat.reshape(-1, 1))
    out = (C.c_double * 2)()
    check(lib().ss_auroc_auprc_mat(ctx.h, dy.h, ds.h, out))
    return float(out[0]), float(out[1])


def AuROC(y, yhat) -> float:
    """reference src/performance.jl:49-63."""
    return _auc_pair(y, yhat)[0]


def AuPRC(y, yhat) -> float:
    """reference src/performance.jl:74-89."""
    return _auc_pair(y, yhat)[1]


def _atl(y, yhat, L: int) -> Tuple[float, float]:
    """y, yhat: (groups, n) matrices; returns (mean recall@L, mean precision@L)."""
    ctx = Context.default()
    dy, ds = DMat.from_host(ctx, y), DMat.from_host(ctx, yhat)
    out = (C.c_double * 2)()
    check(lib().ss_atl(ctx.h, dy.h, ds.h, int(L), out))
    return float(out[0]), float(out[1])


def _atl_dispatch(which: int, y, yhat, *rest):
    y = np.asarray(y, dtype=np.float64).ravel()
    yhat = np.asarray(yhat, dtype=np.float64).ravel()
    if len(rest) >= 1 and not np.isscalar(rest[0]):
        grouping = np.asarray(rest[0]).ravel()
        L = int(rest[1]) if len(rest) > 1 else 20
        assert len(yhat) == len(grouping), "Number of groups must match number of predictions"
        assert len(y) == len(grouping), "Number of groups must match number of labels"
        assert len(y) == len(yhat), "Number of predictions must match number of labels"
        assert L > 0, "Please use a list length greater than 0 (L > 0)"
        order, seen = [], set()
        for gname in grouping.tolist():
            if gname not in seen:
                seen.add(gname)
                order.append(gname)
        masks = [grouping == gname for gname in order]
        sizes = {int(m.sum()) for m in masks}
        if len(sizes) == 1:  # rectangular: one launch for all groups
            Ym = np.stack([y[m] for m in masks])
            Sm = np.stack([yhat[m] for m in masks])
            return _atl(Ym, Sm, L)[which]
        vals = [_atl(y[m].reshape(1, -1), yhat[m].reshape(1, -1), L)[which] for m in masks]
        return float(np.sum(vals) / len(vals))
    L = int(rest[0]) if rest else 20
    assert L > 0, "Please use a list length greater than 0 (L > 0)"
    assert len(y) == len(yhat), "Number of predictions and labels don't match"
    return _atl(y.reshape(1, -1), yhat.reshape(1, -1), L)[which]


def recallatL(y, yhat, *rest) -> float:
    """`recallatL(y, yhat, L=20)` / `recallatL(y, yhat, grouping, L=20)` (reference
    src/performance.jl:308-328, 341-357)."""
    return _atl_dispatch(0, y, yhat, *rest)


def precisionatL(y, yhat, *rest) -> float:
    """`precisionatL(y, yhat, L=20)` / `precisionatL(y, yhat, grouping, L=20)` (reference
    src/performance.jl:370-385, 398-414)."""
    return _atl_dispatch(1, y, yhat, *rest)


def validity_ratio(yhat) -> float:
    """reference src/performance.jl:558-560: `sum(!iszero, yhat) / length(yhat)`."""
    yhat = np.asarray(yhat, dtype=np.float64).ravel()
    return int(k(yhat)) / yhat.size


# ---- confusion-matrix scalars: O(1) host arithmetic in the reference as well ------

_FLOATMIN = 2.2250738585072014e-308


def _chk(tn, fp, fn, tp):
    assert tn + fp + fn + tp > 0, "Confusion matrix sums zero!"


def f1score(tn, fp, fn, tp):
    _chk(tn, fp, fn, tp)
    den = tp + 0.5 * (fp + fn)
    return float("nan") if den == 0 else tp / den


def mcc(*a):
    """`mcc(a, b, eps=floatmin)` (reference src/performance.jl:150-152) or `mcc(tn, fp, fn, tp)`
    (:170-200)."""
    if len(a) in (2, 3):
        x, y = a[0], a[1]
        e = a[2] if len(a) == 3 else _FLOATMIN
        return (x * e - y * e) / np.sqrt((x + y) * (x + e) * (y + e) * (e + e))
    tn, fp, fn, tp = a
    _chk(tn, fp, fn, tp)
    p_pred, n_pred, p_act, n_act = tp + fp, fn + tn, tp + fn, fp + tn
    if p_pred == 0:
        return mcc(tn, fn)
    if n_pred == 0:
        return mcc(tp, fp)
    if p_act == 0:
        return mcc(tn, fp)
    if n_act == 0:
        return mcc(tp, fn)
    return ((tp * tn) - (fp * fn)) / np.sqrt(float(p_pred) * n_pred * p_act * n_act)


def accuracy(tn, fp, fn, tp):
    _chk(tn, fp, fn, tp)
    den = (tp + tn) + (fp + fn)
    return float("nan") if den == 0 else (tp + tn) / den


def balancedaccuracy(tn, fp, fn, tp):
    _chk(tn, fp, fn, tp)
    with np.errstate(divide="ignore", invalid="ignore"):
        return float((np.float64(tp) / np.float64(tp + fn) + np.float64(tn) / np.float64(tn + fp)) / 2)


def recall(tn, fp, fn, tp):
    _chk(tn, fp, fn, tp)
    return float("nan") if tp + fn == 0 else tp / (tp + fn)


def precision(tn, fp, fn, tp):
    _chk(tn, fp, fn, tp)
    return float("nan") if tp + fp == 0 else tp / (tp + fp)


# ------------------------------------------------------------------------------------------------
# Drivers for the loops the reference leaves to user code (docs/src/api.md:17-21):
# k-fold de-novo cross-validation and the alpha sweep.  Everything stays on the GPU between the
# upload of (DD, DT) and the download of the predictions / metrics.
# ------------------------------------------------------------------------------------------------


def _names_only(rows, cols) -> NamedArray:
    """A NamedArray that carries names but no values (index bookkeeping for device-resident data)."""
    n = NamedArray.__new__(NamedArray)
    n.array = np.empty((len(rows), 0), dtype=np.bool_)
    n._names = [[str(r) for r in rows], [str(c) for c in cols]]
    return n


def _fold_indices(X: NamedArray, y: NamedArray, queries: Sequence[str]):
    """Index lists of construct(y, X, queries) (reference src/core.jl:152-154)."""
    qset = set(queries)
    xr, xc = X.names(1), X.names(2)
    features = [f for f in xc if f.lstrip("f") not in qset]
    sources = [d for d in xr if d not in qset]
    _assert_names_differ(features, sources, "Source and Features nodes have the same names!")
    return (X.index_of(queries, 1), X.index_of(sources, 1), X.index_of(features, 2), y.index_of(sources, 1),
            y.index_of(queries, 1))


class _FoldIndexer:
    """`_fold_indices` for many query sets over the same (X, y): the name tables are built once and a fold costs a few
    NumPy mask operations (a 10-fold CV of an Enzyme-sized data set spent a third of its time in the per-fold name
    lookups).  Falls back to `_fold_indices` when names repeat (first-occurrence semantics of `index_of`)."""

    def __init__(self, X: NamedArray, y: NamedArray):
        self.X, self.y = X, y
        xr, xc, yr = X.names(1), X.names(2), y.names(1)
        self.unique = len(set(xr)) == len(xr) and len(set(xc)) == len(xc) and len(set(yr)) == len(yr)
        if not self.unique:
            return
        self.xr = np.array(xr, dtype=object)
        self.xc = np.array(xc, dtype=object)
        self.xrow = {n: i for i, n in enumerate(xr)}
        self.yrow = {n: i for i, n in enumerate(yr)}
        self.cols_of = {}  # stripped feature name -> its columns (construct strips ALL leading 'f', src/core.jl:152)
        for j, f in enumerate(xc):
            self.cols_of.setdefault(f.lstrip("f"), []).append(j)
        self.y_of_xrow = np.array([self.yrow.get(n, -1) for n in xr], dtype=np.int64)
        self.row_order = np.array(sorted(range(len(xr)), key=xr.__getitem__), dtype=np.int64)  # sort(sources) = filter
        self.col_order = np.array(sorted(range(len(xc)), key=xc.__getitem__), dtype=np.int64)  # of the sorted names

    def __call__(self, queries: Sequence[str]):
        if not self.unique or len(set(queries)) != len(queries):
            return _fold_indices(self.X, self.y, queries)
        try:
            qi = np.array([self.xrow[q] for q in queries], dtype=np.int32)
            yqi = np.array([self.yrow[q] for q in queries], dtype=np.int32)
        except KeyError:
            return _fold_indices(self.X, self.y, queries)  # raises the reference-style KeyError
        rows = np.ones(len(self.xr), dtype=bool)
        rows[qi] = False
        cols = np.ones(len(self.xc), dtype=bool)
        for q in queries:
            for j in self.cols_of.get(q, ()):
                cols[j] = False
        si, fi = np.flatnonzero(rows), np.flatnonzero(cols)
        ysi = self.y_of_xrow[si]
        if (ysi < 0).any():
            return _fold_indices(self.X, self.y, queries)
        # reference: @assert all(sort(features) .!= sort(sources))  (element-wise; src/core.jl:156)
        sf = self.xc[self.col_order[cols[self.col_order]]]
        ss_ = self.xr[self.row_order[rows[self.row_order]]]
        if len(sf) != len(ss_):
            raise ValueError("DimensionMismatch: arrays could not be broadcast to a common size; "
                             f"got a dimension with lengths {len(sf)} and {len(ss_)}")
        assert bool(np.all(sf != ss_)), "Source and Features nodes have the same names!"
        return qi, si.astype(np.int32), fi.astype(np.int32), ysi.astype(np.int32), yqi


def cross_validate(DT: NamedArray, DD: NamedArray, alpha: float, weighted: bool = True, k_: int = 10,
                   seed: int = 1, L: int = 20, folds: Optional[List[List[str]]] = None, rank: int = 0,
                   world: int = 1, timing: Optional[dict] = None) -> dict:
    """k-fold de-novo cross-validation: `split` -> `featurize` -> per fold `construct`, `predict`,
    `clean!` -> AuROC / AuPRC over all (query, target) pairs and mean recall@L / precision@L per
    query.  Returns the predictions with rows in fold order.

    Folds are independent (outer level): with `world` > 1 rank r evaluates folds[r::world] only and
    returns {"folds", "yhat", "y"} for those folds (no metrics); `merge_cross_validation` joins the parts
    gathered on the host and computes the metrics over all folds."""
    ctx = Context.default()
    assert DT.size(1) == DD.size(1), "Labels and features have different number of source nodes"
    if folds is None:
        folds = split(DT, k_, seed=seed)
    all_folds = folds
    if world > 1:
        folds = [f for i, f in enumerate(all_folds) if i % world == rank]
    order = [q for f in folds for q in f]
    N, nt = DT.size(1), DT.size(2)
    dX = DMat.from_host(ctx, DD.array)
    check(lib().ss_featurize(ctx.h, dX.h, float(alpha), int(bool(weighted)), dX.h))  # featurize! on device
    Xn = _names_only(DD.names(1), ["f" + c for c in DD.names(2)])  # the values live on the GPU
    dy = DMat.from_host(ctx, DT.array)
    Rall = DMat(ctx, len(order), nt)
    # index lists of every fold (construct's name filtering, src/core.jl:152-154), then ONE library call that queues
    # gather -> degrees -> spread -> T -> R (+ clean!) for all folds and synchronises once (ss_predict_query_folds)
    q_ptr, s_ptr, f_ptr = [0], [0], [0]
    q_all, s_all, ys_all, f_all = [], [], [], []
    indexer = _FoldIndexer(Xn, DT)
    for queries in folds:
        if queries:
            qi, si, fi, ysi, _ = indexer(queries)
            q_all.append(qi), s_all.append(si), ys_all.append(ysi), f_all.append(fi)
        q_ptr.append(q_ptr[-1] + (len(qi) if queries else 0))
        s_ptr.append(s_ptr[-1] + (len(si) if queries else 0))
        f_ptr.append(f_ptr[-1] + (len(fi) if queries else 0))

    def _i32(parts):
        return np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros(0), dtype=np.int32)

    qa, sa, ysa, fa = _i32(q_all), _i32(s_all), _i32(ys_all), _i32(f_all)
    qp, sp, fp = (np.asarray(v, dtype=np.int32) for v in (q_ptr, s_ptr, f_ptr))
    if len(order):
        import time as _time
        l0, t0 = ctx.launch_count(), _time.perf_counter()
        check(lib().ss_predict_query_folds(ctx.h, dX.h, dy.h, len(folds), qp.ctypes.data, qa.ctypes.data, sp.ctypes.data,
                                           sa.ctypes.data, ysa.ctypes.data, fp.ctypes.data, fa.ctypes.data, Rall.h,
                                           SS_PREDICT_CLEAN))
        if timing is not None:  # the call is synchronous: index upload + every fold's degrees / T / R + clean!
            timing["folds_call_ms"] = (_time.perf_counter() - t0) * 1e3
            timing["folds_call_launches"] = ctx.launch_count() - l0
    perm = DIVec.from_host(ctx, DT.index_of(order, 1))
    Yall = DMat(ctx, len(order), nt)
    check(lib().ss_gather(ctx.h, dy.h, perm.h, None, Yall.h))
    if world > 1:  # partial result: metrics need every fold (merge_cross_validation)
        return {"folds": folds, "fold_ids": list(range(rank, len(all_folds), world)),
                "yhat": NamedArray(Rall.to_host(), (order, DT.names(2))),
                "y": NamedArray(Yall.to_host(), (order, DT.names(2)))}
    res = {"folds": folds}
    res.update(_cv_metrics(ctx, Yall, Rall, nt, L))
    res["yhat"] = NamedArray(Rall.to_host(), (order, DT.names(2)))
    res["y"] = NamedArray(Yall.to_host(), (order, DT.names(2)))
    return res


def _cv_metrics(ctx, Yall: "DMat", Rall: "DMat", nt: int, L: int) -> dict:
    auc = (C.c_double * 2)()
    check(lib().ss_auroc_auprc_mat(ctx.h, Yall.h, Rall.h, auc))
    res = {"AuROC": float(auc[0]), "AuPRC": float(auc[1])}
    if nt > L:
        atl = (C.c_double * 2)()
        check(lib().ss_atl(ctx.h, Yall.h, Rall.h, int(L), atl))
        res["recallatL"], res["precisionatL"] = float(atl[0]), float(atl[1])
    return res


def merge_cross_validation(parts: Sequence[dict], L: int = 20) -> dict:
    """Joins the per-rank results of `cross_validate(..., rank=r, world=n)` (gathered on the host, e.g. with
    `torch.distributed.all_gather_object`) back into fold order and computes the metrics of the whole
    cross-validation on the device, exactly as the single-GPU call does."""
    ctx = Context.default()
    by_fold = {}
    for part in parts:
        off = 0
        for fid, names in zip(part["fold_ids"], part["folds"]):
            by_fold[fid] = (names, part["yhat"].array[off:off + len(names)], part["y"].array[off:off + len(names)])
            off += len(names)
    ids = sorted(by_fold)
    folds = [by_fold[i][0] for i in ids]
    order = [q for f in folds for q in f]
    cols = parts[0]["yhat"].names(2)
    yhat = np.concatenate([by_fold[i][1] for i in ids], axis=0) if ids else np.zeros((0, len(cols)))
    y = np.concatenate([by_fold[i][2] for i in ids], axis=0) if ids else np.zeros((0, len(cols)))
    res = {"folds": folds}
    res.update(_cv_metrics(ctx, DMat.from_host(ctx, y), DMat.from_host(ctx, yhat), len(cols), L))
    res["yhat"] = NamedArray(yhat, (order, cols))
    res["y"] = NamedArray(y, (order, cols))
    return res


def alpha_sweep(DT: NamedArray, DD: NamedArray, queries: Sequence[str], alphas: Sequence[float],
                weighted: bool = True, L: int = 20, rank: int = 0, world: int = 1, layout: str = "auto",
                timing: Optional[dict] = None) -> List[dict]:
    """SimSpread's alpha sweep (BASELINE config 3): for every cutoff alpha, featurize -> construct
    -> predict -> clean! -> metrics on the same query set.  alpha points are independent, so with
    `world` > 1 rank r evaluates alphas[r::world] (no communication).

    The raw similarity blocks S[queries, features] and S[sources, features] are extracted once; per alpha
    they are thresholded either into dense feature blocks (DMMA chain) or straight into CSR (row-split
    sparse chain) -- `layout="auto"` picks the sparse chain below SPARSE_DENSITY_THRESHOLD, as `predict`."""
    assert layout in ("auto", "dense", "sparse")
    import time as _time
    ctx = Context.default()
    t_start = _time.perf_counter()
    queries = [str(q) for q in queries]
    dS = DMat.from_host(ctx, DD.array)
    dy = DMat.from_host(ctx, DT.array)
    Xn = _names_only(DD.names(1), ["f" + c for c in DD.names(2)])
    qi, si, fi, ysi, yqi = _fold_indices(Xn, DT, queries)
    nt = DT.size(2)
    dqi, dsi, dfi, dysi, dyqi = (DIVec.from_host(ctx, a) for a in (qi, si, fi, ysi, yqi))
    Sq, Ss = DMat(ctx, len(qi), len(fi)), DMat(ctx, len(si), len(fi))  # raw similarities of the two blocks
    check(lib().ss_gather(ctx.h, dS.h, dqi.h, dfi.h, Sq.h))
    check(lib().ss_gather(ctx.h, dS.h, dsi.h, dfi.h, Ss.h))
    del dS
    Xq, Xs = DMat(ctx, len(qi), len(fi)), DMat(ctx, len(si), len(fi))
    Y, Yq, R = DMat(ctx, len(si), nt), DMat(ctx, len(qi), nt), DMat(ctx, len(qi), nt)
    check(lib().ss_gather(ctx.h, dy.h, dysi.h, None, Y.h))
    check(lib().ss_gather(ctx.h, dy.h, dyqi.h, None, Yq.h))
    kq = DIVec(ctx, len(qi))
    w = int(bool(weighted))
    out = []
    ctx.sync()
    t_setup = _time.perf_counter()
    kx = DIVec(ctx, len(qi))
    for a in list(alphas)[rank::world]:
        use_sparse, have_xq = False, False
        if layout != "dense" and len(qi) and len(si) and len(fi):
            maybe = layout == "sparse"
            if not maybe:  # density of the query block from its thresholded form (needed by the dense chain anyway):
                check(lib().ss_featurize(ctx.h, Sq.h, float(a), w, Xq.h))  # no CSR is built just to be thrown away
                check(lib().ss_k_rows(ctx.h, Xq.h, kx.h))
                have_xq = True
                maybe = kx.to_host().astype(np.int64).sum() < SPARSE_DENSITY_THRESHOLD * len(qi) * len(fi)
            if maybe:
                cq = DCsr.from_dense(ctx, Sq, float(a), bool(weighted))
                cs = DCsr.from_dense(ctx, Ss, float(a), bool(weighted), by_columns=True)
                use_sparse = layout == "sparse" or (cq.density < SPARSE_DENSITY_THRESHOLD
                                                    and cs.density < SPARSE_DENSITY_THRESHOLD)
        if use_sparse:
            check(lib().ss_predict_query_csr(ctx.h, cq.h, cs.h, Y.h, R.h, SS_PREDICT_CLEAN, None))
        else:
            if not have_xq:
                check(lib().ss_featurize(ctx.h, Sq.h, float(a), w, Xq.h))
            check(lib().ss_featurize(ctx.h, Ss.h, float(a), w, Xs.h))
            check(lib().ss_predict_query(ctx.h, Xq.h, Xs.h, Y.h, R.h, SS_PREDICT_CLEAN, None))
        auc, atl = (C.c_double * 2)(), (C.c_double * 2)()
        check(lib().ss_auroc_auprc_mat(ctx.h, Yq.h, R.h, auc))
        rec = {"alpha": float(a), "AuROC": float(auc[0]), "AuPRC": float(auc[1]), "layout": "sparse" if use_sparse else "dense"}
        if nt > L:
            check(lib().ss_atl(ctx.h, Yq.h, R.h, int(L), atl))
            rec["recallatL"], rec["precisionatL"] = float(atl[0]), float(atl[1])
        check(lib().ss_k_rows(ctx.h, R.h, kq.h))
        rec["validity_ratio"] = float(kq.to_host().astype(np.int64).sum() / (len(qi) * nt))
        out.append(rec)
    if timing is not None:  # setup = upload of S / y + block extraction (once per rank); sweep = the alpha points
        timing["setup_s"] = t_setup - t_start
        timing["sweep_s"] = _time.perf_counter() - t_setup
    return out


# ------------------------------------------------------------------------------------------------
# the rest of the metric family on the shared device sort, and `save`
# ------------------------------------------------------------------------------------------------


def BEDROC(y, yhat, rev: bool = True, alpha: float = 20.0) -> float:
    """reference src/performance.jl:22-38 (positives are `y .== 1`, ranks from the stable
    `sortperm(yhat; rev=rev)`)."""
    ctx = Context.default()
    y = np.asarray(y).ravel()
    yhat = np.asarray(yhat, dtype=np.float64).ravel()
    assert len(y) == len(yhat), "The number of scores must be equal to the number of labels"
    dy = DMat.from_host(ctx, (y == 1).astype(np.float64).reshape(-1, 1))
    ds = DMat.from_host(ctx, yhat.reshape(-1, 1))
    out = C.c_double()
    check(lib().ss_bedroc(ctx.h, dy.h, ds.h, int(bool(rev)), float(alpha), C.byref(out)))
    return float(out.value)


def _metric_id(metric) -> int:
    table = {f1score: 0, mcc: 1, accuracy: 2, balancedaccuracy: 3, recall: 4, precision: 5}
    if metric not in table:
        raise TypeError("metric must be one of f1score, mcc, accuracy, balancedaccuracy, recall, precision")
    return table[metric]


def _sweep(y, yhat, metric):
    ctx = Context.default()
    y = np.asarray(y).ravel()
    yhat = np.asarray(yhat, dtype=np.float64).ravel()
    assert len(y) == len(yhat), "The number of scores must be equal to the number of labels"
    dy = DMat.from_host(ctx, (y != 0).astype(np.float64).reshape(-1, 1))
    ds = DMat.from_host(ctx, yhat.reshape(-1, 1))
    out = (C.c_double * 4)()
    check(lib().ss_threshold_sweep(ctx.h, dy.h, ds.h, _metric_id(metric), out))
    return float(out[0]), float(out[1]), float(out[2])


def maxperformance(y, yhat, metric) -> float:
    """reference src/performance.jl:425-448: maximum of `metric` over the confusion matrices of all
    unique-score thresholds."""
    return _sweep(y, yhat, metric)[0]


def meanperformance(y, yhat, metric) -> float:
    """reference src/performance.jl:459-489."""
    return _sweep(y, yhat, metric)[1]


def meanstdperformance(y, yhat, metric) -> Tuple[float, float]:
    """reference src/performance.jl:500-531 (`mean_and_std`, corrected sample std)."""
    _, m, sd = _sweep(y, yhat, metric)
    return m, sd


def _jl_string(x) -> str:
    """Julia `string(x)` for the numbers `save` / `writedlm` write (Int or Float64: shortest round-trip digits,
    fixed notation for decimal exponents -4..5, `d.ddde±x` otherwise -- Base.Ryu.writeshortest)."""
    if isinstance(x, (bool, np.bool_)):
        return "true" if x else "false"
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    if x == 0:
        return "-0.0" if np.signbit(x) else "0.0"
    from decimal import Decimal
    sign, digits, exp = Decimal(repr(x)).as_tuple()
    digits = "".join(map(str, digits)).rstrip("0") or "0"
    e10 = len("".join(map(str, Decimal(repr(x)).as_tuple().digits))) + exp - 1  # x = d.ddd * 10^e10
    neg = "-" if sign else ""
    if -5 < e10 < 6:
        if e10 >= 0:
            ip = (digits + "0" * (e10 + 1))[:e10 + 1]
            fp = digits[e10 + 1:] or "0"
        else:
            ip, fp = "0", "0" * (-e10 - 1) + digits
        return f"{neg}{ip}.{fp}"
    return f"{neg}{digits[0]}.{digits[1:] or '0'}e{e10}"


def read_namedmatrix(filepath: str, delimiter: str = " ", valuetype=float, rows: bool = True, cols: bool = True) -> NamedArray:
    """`read_namedmatrix(filepath, delimiter, valuetype; rows, cols)` (reference src/utils.jl:50-53 and
    `_parse_matrix` :24-40): names from the first row / column (or `R#i` / `C#j`), values parsed as Float64,
    rows and columns re-ordered by sorted name (:38).  The value block is parsed by the library's
    multi-threaded host reader (`ss_text_matrix_read`), the names here."""
    assert len(delimiter) == 1, "delimiter must be a single character"
    path = os.fspath(filepath).encode()
    d = ord(delimiter)
    nl, nf = C.c_int64(), C.c_int64()
    check(lib().ss_text_matrix_dims(path, d, C.byref(nl), C.byref(nf)))
    nr, nc = nl.value - int(cols), nf.value - int(rows)
    if nr < 0 or nc < 0:
        raise ValueError(f"{filepath}: no value block")
    vals = np.zeros((nr, nc), order="F")
    check(lib().ss_text_matrix_read(path, d, int(cols), int(rows), vals.ctypes.data, nr, nc, max(nr, 1)))
    row_names = [f"R#{i + 1}" for i in range(nr)]
    col_names = [f"C#{j + 1}" for j in range(nc)]
    if rows or cols:
        with open(filepath, "r", newline="") as f:
            for i, line in enumerate(f):
                line = line.rstrip("\n").rstrip("\r")
                if i == 0 and cols:
                    col_names = line.split(delimiter)[int(rows):]
                    if not rows:
                        break
                    continue
                if not rows:
                    break
                if i - int(cols) < nr:
                    row_names[i - int(cols)] = line.split(delimiter, 1)[0]
    if valuetype is not float and valuetype is not np.float64:
        vals = vals.astype(valuetype)
    ro = sorted(range(nr), key=lambda i: row_names[i])
    co = sorted(range(nc), key=lambda j: col_names[j])
    return NamedArray(vals[np.ix_(ro, co)], ([row_names[i] for i in ro], [col_names[j] for j in co]))


def writedlm(io, x: NamedArray, delimiter: str = "\t") -> None:
    """`writedlm(io, x::NamedMatrix[, delimiter])` (reference src/utils.jl:6-11): the matrix
    `["" names(x, 2)...; names(x, 1) x]` written with Julia's `print` of each cell."""
    close = isinstance(io, (str, os.PathLike))
    f = open(io, "w") if close else io
    try:
        f.write(delimiter.join([""] + [str(c) for c in x.names(2)]) + "\n")
        arr = x.array
        for i, r in enumerate(x.names(1)):
            f.write(delimiter.join([str(r)] + [_jl_string(v) for v in arr[i]]) + "\n")
    finally:
        if close:
            f.close()


def save(filepath: str, *args, delimiter: str = "\t") -> None:
    """`save(filepath, yhat, y; delimiter)` (reference src/core.jl:503-522: the fold column is the
    1-based index of the query) and `save(filepath, fidx, yhat, y; delimiter)` (:542-561).  Rows
    `fold, "source", "target", score, label` are APPENDED ("a+"), as in the reference."""
    if len(args) == 2:
        fidx, (yhat, y) = None, args
    elif len(args) == 3:
        fidx, yhat, y = args
    else:
        raise TypeError("MethodError: no method matching save(...)")
    queries, targets = y.names(1), y.names(2)
    yh = yhat[queries, targets].array
    yy = y.array
    native = (len(delimiter) == 1 and ord(delimiter) < 128 and (fidx is None or isinstance(fidx, (int, np.integer)))
              and all(a.dtype != np.bool_ and (np.issubdtype(a.dtype, np.floating) or np.issubdtype(a.dtype, np.integer)) for a in (yh, yy))
              and len(set(queries)) == len(queries))
    if native:
        # the library formats and writes the rows with all host cores (ss_save_rows; Julia's number format in C++)
        qn = (C.c_char_p * len(queries))(*[str(q).encode() for q in queries])
        tn_ = (C.c_char_p * len(targets))(*[str(t).encode() for t in targets])
        a_h, a_y = np.asfortranarray(yh, dtype=np.float64), np.asfortranarray(yy, dtype=np.float64)
        nbytes = C.c_int64()
        check(lib().ss_save_rows(os.fspath(filepath).encode(), 1, -1 if fidx is None else int(fidx), len(queries), len(targets),
                                 qn, tn_, a_h.ctypes.data, max(1, a_h.shape[0]), int(np.issubdtype(yh.dtype, np.integer)),
                                 a_y.ctypes.data, max(1, a_y.shape[0]), int(np.issubdtype(yy.dtype, np.integer)),
                                 ord(delimiter), C.byref(nbytes)))
        return
    with open(filepath, "a+") as f:  # exotic element types / delimiters: the per-cell path
        for qi, q in enumerate(queries):
            fold = (queries.index(q) + 1) if fidx is None else fidx
            for ti, t in enumerate(targets):
                row = [_jl_string(fold), '"' + q + '"', '"' + t + '"', _jl_string(yh[qi, ti]), _jl_string(yy[qi, ti])]
                f.write(delimiter.join(row) + "\n")


def recommend_topl(y, L: int = 20, weighted: Optional[bool] = None, s_range: Optional[Tuple[int, int]] = None):
    """Top-L targets per source of the classical 2-layer NBI (`predict(construct(y, X), y)` of the
    reference with an empty feature layer, src/core.jl:446-466, ranked as in src/performance.jl:315),
    computed from the sparse graph without materialising the score matrix (BASELINE config 5).
    `y`: NamedArray / dense matrix (sources x targets).  Returns (idx, val): (sources, L) arrays of
    0-based target indices (-1 padding) and scores.  `s_range=(begin, end)` restricts the sources
    (multi-GPU sharding by source rows)."""
    ctx = Context.default()
    arr = y.array if isinstance(y, NamedArray) else np.asarray(y)
    ns, nt = arr.shape
    if weighted is None:
        weighted = bool(np.any((arr != 0) & (arr != 1)))
    d = DMat.from_host(ctx, arr)
    tiny = 5e-324  # keep every positive entry
    cy = DCsr.from_dense(ctx, d, tiny if not weighted else float("-inf"), weighted)
    cyt = DCsr.from_dense(ctx, d, tiny if not weighted else float("-inf"), weighted, by_columns=True)
    idx = DIVec(ctx, L * ns)
    val = DMat(ctx, L, ns)
    b, e = s_range if s_range is not None else (0, ns)
    check(lib().ss_recommend_topl(ctx.h, cy.h, cyt.h, int(L), int(b), int(e), idx.h, val.h))
    v = val.to_host().T
    return idx.to_host().reshape(ns, L), v
