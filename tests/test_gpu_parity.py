"""`-m gpu` parity tests: the CUDA path (through the C ABI / the host mirror of the reference API)
against the CPU oracle on the same seeded inputs, and against the reference's golden vectors.

Bars (BASELINE.json north_star): bit-exact for thresholds, CSR indices, degrees and top-L order;
<= 1e-12 relative error for FP64 scores."""
import ctypes as C
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-12  # FP64 score tolerance stated by north_star


@pytest.fixture(scope="module")
def ss():
    import simspread_b200 as m
    m.build()
    m.Context.default()  # raises without a B200: no CPU fallback
    return m


@pytest.fixture(scope="module")
def o():
    from oracle import simspread_oracle
    return simspread_oracle


def relerr(got, want):
    """Element-wise relative error (all operands of the path are non-negative, so there is no
    cancellation); entries whose expected value is exactly 0 must be exactly 0."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    zero = want == 0
    if np.any(got[zero] != 0):
        return float("inf")
    if zero.all():
        return 0.0
    return float(np.max(np.abs(got[~zero] - want[~zero]) / np.abs(want[~zero])))


# ---------------------------------------------------------------------------------------------
# golden vectors of the reference test-suite, through the API mirror
# ---------------------------------------------------------------------------------------------


def test_kat_k(ss, kats):
    M = np.array(kats["k"]["M"], dtype=float)
    assert ss.k(1, M) == 0
    assert ss.k(M[0, :]) == 0
    assert ss.k(M).ravel().tolist() == kats["k"]["expect"]
    assert ss.k(M).shape == (4, 1)
    assert ss.k(np.array([0.0, -0.0, np.nan, 1.0])) == 2


def test_kat_cutoff_and_featurize(ss, kats):
    c = kats["cutoff"]
    x, y, z = c["x"], np.array(c["y"]).reshape(-1, 1), np.array(c["z"])
    for case in c["cases"]:
        a = case["alpha"]
        assert ss.cutoff(x, a, False) == case["x_bin"]
        assert ss.cutoff(x, a, True) == case["x_w"]
        yw = y.ravel() if case["y_w"] == "y" else np.array(case["y_w"], dtype=float)
        zw = z if case["z_w"] == "z" else np.array(case["z_w"], dtype=float)
        assert np.array_equal(ss.cutoff(y, a, False).ravel(), np.array(case["y_bin"], dtype=float))
        assert np.array_equal(ss.cutoff(y, a, True).ravel(), yw)
        assert np.array_equal(ss.cutoff(z, a, False), np.array(case["z_bin"], dtype=float))
        assert np.array_equal(ss.cutoff(z, a, True), zw)
    zz = z.copy()
    ss.cutoff_(zz, 0.5, False)
    assert np.array_equal(zz, z)  # cutoff! does not mutate (reference quirk)
    f = kats["featurize"]
    M0 = ss.NamedArray(np.array(f["M0"]), (f["names"], f["names"]))
    b, w = ss.featurize(M0, f["alpha"], False), ss.featurize(M0, f["alpha"], True)
    assert np.array_equal(b.array, np.array(f["bin"], dtype=float)) and b.names(2) == f["colnames"]
    assert np.array_equal(w.array, np.array(f["w"], dtype=float))
    assert M0.names(2) == f["names"]  # featurize is not in place
    ss.featurize_(M0, f["alpha"], False)
    assert np.array_equal(M0.array, np.array(f["bin"], dtype=float)) and M0.names(2) == f["colnames"]


def test_kat_construct(ss, kats, o):
    c = kats["construct"]
    X = ss.NamedArray(np.array(c["X"], dtype=float), (c["xrows"], c["xcols"]))
    y = ss.NamedArray(np.array(c["y"], dtype=float), (c["xrows"], c["ycols"]))
    A, B = ss.construct(y, X, c["queries"])
    assert A.names(1) == A.names(2) == c["names"]
    assert B.names(1) == B.names(2) == c["names"]
    Ao, Bo, _ = o.construct_queries(y.array, (c["xrows"], c["ycols"]), X.array, (c["xrows"], c["xcols"]),
                                    c["queries"])
    assert np.array_equal(A.array, Ao) and np.array_equal(B.array, Bo)


def test_kat_spread(ss, kats):
    s = kats["spread"]
    W = ss.spread(np.array(s["M"], dtype=float))
    assert np.array_equal(W, np.array([[1, 0, 0], [.5, .5, 0], [1 / 3, 1 / 3, 1 / 3]]))
    assert np.array_equal(ss.spread(np.zeros((3, 3))), np.zeros((3, 3)))  # 0/0 -> NaN -> 0


def test_kat_predict_exact(ss, kats):
    p = kats["predict"]
    A = ss.NamedArray(np.array(p["A"], dtype=float), (p["names"], p["names"]))
    B = ss.NamedArray(np.array(p["B"], dtype=float), (p["names"], p["names"]))
    y = ss.NamedArray(np.array([[0., 1.]]), (p["rows"], p["cols"]))
    want = np.array(p["yhat"])
    assert np.array_equal(ss.predict(A, B, y).array, want)      # exact ==, as the reference test
    assert np.array_equal(ss.predict((A, B), y).array, want)
    # same graph through construct (block-reduced path)
    Xt = ss.NamedArray(A.array[0:1, 4:7], (["q1"], ["f1", "f2", "f3"]))
    Xs = ss.NamedArray(A.array[1:4, 4:7], (["s1", "s2", "s3"], ["f1", "f2", "f3"]))
    ys = ss.NamedArray(A.array[1:4, 7:9], (["s1", "s2", "s3"], ["t1", "t2"]))
    G = ss.construct(ys, y, Xs, Xt)
    assert np.array_equal(G[0].array, A.array) and np.array_equal(G[1].array, B.array)
    assert np.array_equal(ss.predict(G, y).array, want)


def test_kat_clean(ss, kats):
    c = kats["clean"]
    A = ss.NamedArray(np.array(c["A"], dtype=float), (c["names"], c["names"]))
    yhat = ss.NamedArray(np.array(c["yhat"], dtype=float), (["q1"], c["targets"]))
    y = ss.NamedArray(np.array([[0., 1.]]), (["q1"], c["targets"]))
    ss.clean_(yhat, A, y)
    assert np.array_equal(yhat.array, np.array(c["expect"], dtype=float))


def test_kat_atL(ss, kats):
    a = kats["atL"]
    for L, v in a["recall"].items():
        assert ss.recallatL(a["y"], a["yhat"], a["grouping"], int(L)) == pytest.approx(v, rel=1e-15)
    for L, v in a["precision"].items():
        assert ss.precisionatL(a["y"], a["yhat"], a["grouping"], int(L)) == pytest.approx(v, rel=1e-15)
    with pytest.raises(AssertionError, match="Number of labels is less than length"):
        ss.recallatL(a["y"], a["yhat"], 10)
    assert math.isnan(ss.recallatL([0, 0, 1, 0], [1, 2, 3, 4], [1, 1, 2, 2], 1))


# ---------------------------------------------------------------------------------------------
# kernels against the oracle
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("rows,cols", [(1, 1), (7, 5), (445, 445), (1000, 333), (2049, 130), (3000, 1500)])
@pytest.mark.parametrize("weighted", [False, True])
def test_featurize_dense_and_csr_bit_exact(ss, o, rows, cols, weighted):
    rng = np.random.default_rng(rows * 1000 + cols)
    S = np.round(rng.random((rows, cols)), 3)
    S.flat[rng.integers(0, S.size, size=max(1, S.size // 50))] = np.nan
    S.flat[rng.integers(0, S.size, size=max(1, S.size // 50))] = -0.0
    S.flat[rng.integers(0, S.size, size=max(1, S.size // 50))] = 0.35  # exactly alpha: kept (>=)
    from simspread_b200._lib import check
    ctx = ss.Context.default()
    d = ss.DMat.from_host(ctx, S)
    for alpha in (0.35, -0.01, 0.0, 1.01, 0.97):
        want = o.cutoff(S, alpha, weighted)
        got = ss.cutoff(S, alpha, weighted)
        assert np.array_equal(got, want)
        assert np.array_equal(np.signbit(got), np.signbit(want))
        wr, wc = np.nonzero(want != 0)  # row-major order == CSR order
        # CSR: count + keep-mask with lanes along rows, device-wide scan of the (row, segment) counts, then either
        # fill (mask replay; the default below 10 % density) or the transposing tiled fill -- both forced here
        for fill in ("replay", "tiled", None):
            if fill:
                os.environ["SS_CSR_FILL"] = fill
            try:
                h = C.c_void_p()
                check(ss.lib().ss_featurize_csr(ctx.h, d.h, alpha, int(weighted), C.byref(h)))
            finally:
                os.environ.pop("SS_CSR_FILL", None)
            r, c, nnz, hv = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
            check(ss.lib().ss_csr_info(h, C.byref(r), C.byref(c), C.byref(nnz), C.byref(hv)))
            rp = np.empty(rows + 1, np.int32)
            ci = np.empty(max(nnz.value, 1), np.int32)
            va = np.empty(max(nnz.value, 1), np.float64)
            check(ss.lib().ss_csr_download(ctx.h, h, rp.ctypes.data, ci.ctypes.data, va.ctypes.data if hv.value else None))
            ss.lib().ss_csr_destroy(h)
            assert nnz.value == len(wr) and bool(hv.value) == weighted
            assert np.array_equal(rp, np.concatenate(([0], np.cumsum(np.bincount(wr, minlength=rows)))).astype(np.int32))
            assert np.array_equal(ci[:nnz.value], wc.astype(np.int32))
            if weighted:
                assert np.array_equal(va[:nnz.value], want[wr, wc])


@pytest.mark.parametrize("ns,nf,nt", [(1, 1, 1), (13, 13, 5), (401, 401, 664), (1500, 700, 129)])
def test_degrees_exact(ss, o, ns, nf, nt):
    rng = np.random.default_rng(ns + nf + nt)
    Xs = o.cutoff(np.round(rng.random((ns, nf)), 3), 0.6, True)
    Xs.flat[rng.integers(0, Xs.size, size=max(1, Xs.size // 40))] = np.nan
    Xs.flat[rng.integers(0, Xs.size, size=max(1, Xs.size // 40))] = -0.0
    Y = (rng.random((ns, nt)) < 0.05).astype(float)
    ks, kf, kt = o.degrees_blocks(Xs, Y)
    ctx = ss.Context.default()
    dX, dY = ss.DMat.from_host(ctx, Xs), ss.DMat.from_host(ctx, Y)
    vs, vf, vt = ss.DIVec(ctx, ns), ss.DIVec(ctx, nf), ss.DIVec(ctx, nt)
    from simspread_b200._lib import check
    check(ss.lib().ss_degrees(ctx.h, dX.h, dY.h, vs.h, vf.h, vt.h))
    assert np.array_equal(vs.to_host(), ks) and np.array_equal(vf.to_host(), kf) and np.array_equal(vt.to_host(), kt)
    assert np.array_equal(ss.k(Xs).ravel(), o.k_mat(Xs).ravel())
    # spread on the same matrix: W = G ./ k(G), NaN/Inf -> 0 -- element-wise identical to IEEE division
    G = np.where(np.isnan(Xs), 2.5, Xs)
    assert np.array_equal(ss.spread(G), o.spread(G))


@pytest.mark.parametrize("rows,cols", [(1, 1), (3, 70), (445, 664), (1000, 33), (2100, 2300)])
def test_row_major_upload_is_transposed_on_the_device(ss, rows, cols):
    """NumPy's default order goes up as it lies and is transposed by `transpose_kernel` (ss_mat_upload_rowmajor);
    the device image must equal the column-major upload of the same matrix (2100 x 2300 = 38 MB takes the staged
    pageable-memory copy)."""
    rng = np.random.default_rng(rows + cols)
    a = rng.standard_normal((rows, cols))
    assert a.flags.c_contiguous
    ctx = ss.Context.default()
    from simspread_b200 import host as _host
    old_min, _host._ROWMAJOR_UPLOAD_MIN_BYTES = _host._ROWMAJOR_UPLOAD_MIN_BYTES, 0  # small arrays too
    try:
        got = ss.DMat.from_host(ctx, a).to_host()
    finally:
        _host._ROWMAJOR_UPLOAD_MIN_BYTES = old_min
    want = ss.DMat.from_host(ctx, np.asfortranarray(a)).to_host()
    assert np.array_equal(got, a) and np.array_equal(want, a)
    # a row-major view with a pitch (every second column block of a wider array) takes the generic path
    wide = rng.standard_normal((rows, 2 * cols))
    assert np.array_equal(ss.DMat.from_host(ctx, wide[:, :cols]).to_host(), wide[:, :cols])


GEMM_SHAPES = [(1, 1, 1), (5, 3, 2), (128, 128, 16), (129, 127, 17), (300, 200, 100), (45, 664, 400),
               (1000, 37, 555), (64, 1500, 1031)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("op", ["N", "T"])
def test_gemm_f64_against_numpy(ss, M, N, K, op):
    from simspread_b200._lib import SS_OP_N, SS_OP_T, check
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    ctx = ss.Context.default()
    for integer in (True, False):
        if integer:  # sums exactly representable -> bit-exact regardless of summation order
            A = rng.integers(-3, 4, size=(M, K)).astype(float)
            B = rng.integers(-3, 4, size=(K, N)).astype(float)
        else:
            A, B = rng.standard_normal((M, K)), rng.standard_normal((K, N))
        dA = ss.DMat.from_host(ctx, A if op == "N" else A.T)
        dB, dC = ss.DMat.from_host(ctx, B), ss.DMat(ctx, M, N)
        check(ss.lib().ss_gemm_f64(ctx.h, SS_OP_N if op == "N" else SS_OP_T, dA.h, dB.h, dC.h, None, None))
        got, want = dC.to_host(), A @ B
        if integer:
            assert np.array_equal(got, want)
        else:
            bound = np.abs(A) @ np.abs(B)  # forward error bound scale of a K-term dot product
            assert np.max(np.abs(got - want) / np.maximum(bound, 1e-300)) < 1e-13
        # fused epilogues: row division (k == 0 -> 0) and clean! flag
        div = rng.integers(0, 4, size=M).astype(np.int32)
        flag = rng.integers(0, 2, size=N).astype(np.int32)
        dd, df = ss.DIVec.from_host(ctx, div), ss.DIVec.from_host(ctx, flag)
        check(ss.lib().ss_gemm_f64(ctx.h, SS_OP_N if op == "N" else SS_OP_T, dA.h, dB.h, dC.h, dd.h, df.h))
        with np.errstate(divide="ignore", invalid="ignore"):
            ref = np.where(div[:, None] == 0, 0.0, got / div[:, None].astype(float))
        ref[:, flag == 0] = -99.0
        assert np.array_equal(dC.to_host(), ref)


def _enzyme_like(o, seed=20241):
    """BASELINE config 2 shape: 445 drugs x 664 targets, binary alpha-cutoff features."""
    rng = np.random.default_rng(seed)
    N, Nt = 445, 664
    S = np.round(rng.beta(2, 5, size=(N, N)), 6)
    np.fill_diagonal(S, 1.0)
    Y = (rng.random((N, Nt)) < 0.0099).astype(float)
    return S, Y


@pytest.mark.parametrize("weighted,alpha", [(False, 0.35), (True, 0.2), (True, 0.0)])
def test_predict_cv_fold_against_oracle(ss, o, weighted, alpha):
    S, Yfull = _enzyme_like(o)
    N, Nt = Yfull.shape
    names = [f"D{i:04d}" for i in range(N)]
    tnames = [f"T{j:04d}" for j in range(Nt)]
    DD, DT = ss.NamedArray(S, (names, names)), ss.NamedArray(Yfull, (names, tnames))
    X = ss.featurize(DD, alpha, weighted)
    Xo, xr, xc = o.featurize(S, names, names, alpha, weighted)
    assert np.array_equal(X.array, Xo) and X.names(2) == xc
    folds = ss.split(DT, 10, seed=1)
    assert sorted(sum(folds, [])) == sorted(names)
    for queries in folds[:3]:
        A, B = ss.construct(DT, X, queries)
        ytest = DT[queries, tnames]
        yhat = ss.predict((A, B), ytest)
        Ao, Bo, nn = o.construct_queries(Yfull, (names, tnames), Xo, (xr, xc), queries)
        assert A.names(1) == nn
        want = o.predict_dense(Ao, Bo, nn, queries, tnames)  # literal n x n reference path
        assert yhat.names(1) == queries and yhat.names(2) == tnames
        assert relerr(yhat.array, want) < RTOL
        # clean!: fused flag == separate call == oracle
        want_c = want.copy()
        o.clean(want_c, Ao, nn, tnames)
        fused = ss.predict((A, B), ytest, clean=True)
        ss.clean_(yhat, A, ytest)
        assert relerr(yhat.array, want_c) < RTOL and np.array_equal(yhat.array == -99, want_c == -99)
        assert np.array_equal(fused.array, yhat.array)
        # training rows through the same graph (tutorial usage, src/core.jl:421 with source names)
        some = A.sources[:25]
        ytr = DT[some, tnames]
        assert relerr(ss.predict((A, B), ytr).array, o.predict_dense(Ao, Bo, nn, some, tnames)) < RTOL


def test_predict_three_layer_and_two_layer(ss, o, iris):
    S, Cc, names = iris["S"], iris["C"], iris["names"]
    X = ss.featurize(ss.NamedArray(S, (names, names)), 0.9, True)
    y = ss.NamedArray(Cc, (names, iris["classes"]))
    A = ss.construct(y, X)
    got = ss.predict(A, y)
    Ao, nn = o.construct_3layer(Cc, (names, iris["classes"]), X.array, (names, X.names(2)))
    assert A.names(1) == nn
    want = o.predict_dense_single(Ao, nn, names, iris["classes"])
    assert relerr(got.array, want) < RTOL
    # time-split form on iris (tutorial): 15 queries
    q = names[::10]
    tr = [n for n in names if n not in q]
    Xn = ss.NamedArray(S, (names, names))
    Xtr, Xte = ss.featurize(Xn[tr, tr], 0.9, True), ss.featurize(Xn[q, tr], 0.9, True)
    G = ss.construct(y[tr, iris["classes"]], y[q, iris["classes"]], Xtr, Xte)
    got = ss.predict(G, y[q, iris["classes"]])
    Ao, Bo, nn = o.construct_split(Cc[[names.index(t) for t in tr]], (tr, iris["classes"]),
                                   Cc[[names.index(t) for t in q]], (q, iris["classes"]),
                                   Xtr.array, (tr, Xtr.names(2)), Xte.array, (q, Xte.names(2)))
    want = o.predict_dense(Ao, Bo, nn, q, iris["classes"])
    assert relerr(got.array, want) < RTOL
    assert ss.AuROC(Cc[[names.index(t) for t in q]].ravel() > 0, got.array.ravel()) == pytest.approx(
        o.AuROC(Cc[[names.index(t) for t in q]].ravel() > 0, want.ravel()), rel=1e-12)


def test_c4_replica_one_fiftieth_against_the_literal_path(ss, o):
    """full comparison at a 1/50-scale replica of BASELINE config 4 (2 000 queries, 400 sources = 400
    features, 1 000 targets; dense n = 3 800) against the LITERAL reference path -- construct the n x n matrices,
    W = spread(B), F = A * (W * W), slice, clean! (src/core.jl:148-201, 365-371, 402-423, 478-484)."""
    from simspread_b200._lib import SS_PREDICT_CLEAN, check
    nq, ns, nf, nt = 2000, 400, 400, 1000
    Xq, Xs, Y = o.synth_dense(nq, ns, nf, nt, seed=20244, y_density=0.05, alpha=0.0, weighted=True)
    Y[:, 17] = 0.0
    A = o._assemble4(Xq, Xs, Y)
    B = A.copy()
    B[:nq, :] = 0.0
    B[:, :nq] = 0.0
    names = [str(i) for i in range(nq + ns + nf + nt)]
    rows, cols = names[:nq], names[nq + ns + nf:]
    want = o.predict_dense(A, B, names, rows, cols)
    o.clean(want, A, names, cols)
    ctx = ss.Context.default()
    dq, dx, dy, R = (ss.DMat.from_host(ctx, a) for a in (Xq, Xs, Y, np.zeros((nq, nt))))
    check(ss.lib().ss_predict_query(ctx.h, dq.h, dx.h, dy.h, R.h, SS_PREDICT_CLEAN, None))
    got = R.to_host()
    assert relerr(got, want) < RTOL and np.array_equal(got == -99, want == -99) and (want[:, 17] == -99).all()
    # the same through the host-buffer entry point (what the Julia `predict` wrapper ccalls)
    Rh = np.zeros((nq, nt), order="F")
    Xqf, Xsf, Yf = (np.asfortranarray(a) for a in (Xq, Xs, Y))
    check(ss.lib().ss_predict_query_host(ctx.h, Xqf.ctypes.data, nq, Xsf.ctypes.data, ns, Yf.ctypes.data, ns, nq, ns, nf, nt,
                                         SS_PREDICT_CLEAN, Rh.ctypes.data, nq))
    assert np.array_equal(Rh, got)


@pytest.mark.parametrize("alpha", [0.0, 0.5, 0.97])
def test_c3_full_size_against_block_oracle(ss, o, alpha):
    """BASELINE config 3 at FULL size (5 000 queries, 5 000 sources = features, 2 000 targets, weighted) for three of
    the 21 alpha points: dense end, middle, and the sparse end where predict(layout="auto") takes the row-split chain."""
    nq = ns = 5000
    nt = 2000
    rng = np.random.default_rng(20243)
    S = np.round(rng.random((nq + ns, ns)), 6)
    Yall = (rng.random((nq + ns, nt)) < 0.01).astype(float)
    X = o.cutoff(S, alpha, True)
    Xq, Xs, Y = X[:nq], X[nq:], Yall[nq:]
    want = o.predict_blocks_query(Xq, Xs, Y)
    o.clean_blocks(want, o.degrees_blocks(Xs, Y)[2])
    qn, sn = [f"q{i}" for i in range(nq)], [f"s{i}" for i in range(ns)]
    fn, tn = [f"f{i}" for i in range(ns)], [f"t{i}" for i in range(nt)]
    G = ss.construct((ss.NamedArray(Y, (sn, tn)), ss.NamedArray(Yall[:nq], (qn, tn))),
                     (ss.NamedArray(Xs, (sn, fn)), ss.NamedArray(Xq, (qn, fn))))
    got = ss.predict(G, ss.NamedArray(Yall[:nq], (qn, tn)), clean=True)
    assert G[0].last_layout == ("sparse" if alpha > 0.96 else "dense")  # 3 % dense < SPARSE_DENSITY_THRESHOLD
    assert relerr(got.array, want) < RTOL and np.array_equal(got.array == -99, want == -99)
    yb = Yall[:nq].ravel() > 0
    assert ss.AuROC(yb, got.array.ravel()) == pytest.approx(o.AuROC(yb, want.ravel()), rel=1e-12)


def test_predict_forms_that_are_not_the_block_chain(ss, o, iris):
    """predict() may use the block-reduced chain only where it IS the reference's A * (W * W) with W = spread(B):
    (1) predict(A, y[sources]) on the UNMASKED 4-layer A: W = spread(A) counts the query edges in the feature
    degrees; (2) predict((B, B), yq): the query rows of a masked A are zero; (3) graphs of two different construct()
    calls; (4) query and source rows mixed in one call.  Each against the literal oracle (src/core.jl:402-466)."""
    S, Cc, names, classes = iris["S"], iris["C"], iris["names"], iris["classes"]
    X = ss.featurize(ss.NamedArray(S, (names, names)), 0.9, True)
    y = ss.NamedArray(Cc, (names, classes))
    q = names[::10]
    tr = [n for n in names if n not in q]
    A, B = ss.construct(y, X, q)
    Xo, xr, xc = o.featurize(S, names, names, 0.9, True)
    Ao, Bo, nn = o.construct_queries(Cc, (names, classes), Xo, (xr, xc), q)
    some = tr[3:40:4]
    # (1) W = spread(A), source rows and query rows
    assert relerr(ss.predict(A, y[some, classes]).array, o.predict_dense(Ao, Ao, nn, some, classes)) < RTOL
    assert relerr(ss.predict(A, y[q, classes]).array, o.predict_dense(Ao, Ao, nn, q, classes)) < RTOL
    assert relerr(ss.predict((A, A), y[some, classes]).array, o.predict_dense(Ao, Ao, nn, some, classes)) < RTOL
    # (2) A = B = masked graph: query rows are zero rows, source rows are the block form
    got = ss.predict((B, B), y[q, classes]).array
    assert np.array_equal(got, o.predict_dense(Bo, Bo, nn, q, classes)) and not got.any()
    assert relerr(ss.predict((B, B), y[some, classes]).array, o.predict_dense(Bo, Bo, nn, some, classes)) < RTOL
    # (3) A of one construct() call, B of another (other queries -> other feature / source sets of the same size)
    q2 = names[5::10]
    A2, B2 = ss.construct(y, X, q2)
    A2o, B2o, nn2 = o.construct_queries(Cc, (names, classes), Xo, (xr, xc), q2)
    rows3 = [n for n in q if n in nn2[:len(q2) + len(names) - len(q2)]][:5]
    want3 = (A2o @ (o.spread(Bo) @ o.spread(Bo)))  # literal, by position: the node orders differ, as in the reference
    got3 = ss.predict((A2, B), ss.NamedArray(np.zeros((len(rows3), len(classes))), (rows3, classes)))
    ridx = [nn2.index(r) for r in rows3]
    cidx = [nn2.index(c) for c in classes]
    assert relerr(got3.array, want3[np.ix_(ridx, cidx)]) < RTOL
    # (4) mixed query + source rows with the (A, B) pair: one call, two kernels, no per-row host loop
    mixed = [q[0], some[0], q[3], some[2], some[1], q[1]]
    assert relerr(ss.predict((A, B), y[mixed, classes]).array, o.predict_dense(Ao, Bo, nn, mixed, classes)) < RTOL
    sub = classes[::-1][:2]
    assert relerr(ss.predict((A, B), y[mixed, sub]).array, o.predict_dense(Ao, Bo, nn, mixed, sub)) < RTOL


@pytest.mark.parametrize("nq,ns,nf,nt,slab", [(700, 300, 260, 190, 128), (33, 50, 40, 20, 16), (1, 3, 3, 2, 0)])
def test_predict_query_host_pipelined(ss, o, nq, ns, nf, nt, slab):
    """The reference-facing host-buffer entry point (slabs of query rows streamed through)."""
    from simspread_b200._lib import SS_PREDICT_CLEAN, check
    Xq, Xs, Y = o.synth_dense(nq, ns, nf, nt, seed=5, y_density=0.05, alpha=0.3, weighted=True)
    Y[:, 1 % nt] = 0.0
    want = o.predict_blocks_query(Xq, Xs, Y)
    o.clean_blocks(want, o.degrees_blocks(Xs, Y)[2])
    ctx = ss.Context.default()
    if slab:
        os.environ["SS_SLAB_ROWS"] = str(slab)
    try:
        R = np.full((nq + 3, nt), 7.0, order="F")  # ld > rows: padding must stay untouched
        check(ss.lib().ss_predict_query_host(ctx.h, Xq.ctypes.data, nq, Xs.ctypes.data, ns, Y.ctypes.data, ns,
                                             nq, ns, nf, nt, SS_PREDICT_CLEAN, R.ctypes.data, nq + 3))
    finally:
        os.environ.pop("SS_SLAB_ROWS", None)
    assert relerr(R[:nq], want) < RTOL
    assert np.array_equal(R[:nq] == -99, want == -99)
    assert np.all(R[nq:] == 7.0)


@pytest.mark.parametrize("nq,ns,nf,nt", [(20000, 256, 256, 3500), (40, 24, 24, 10)])
def test_predict_query_fetch_overlapped_download(ss, o, nq, ns, nf, nt):
    """ss_predict_query_fetch (second product in column blocks, each block downloaded while the next is computed) is
    bit-identical to ss_predict_query + download, into pageable memory (staging buffers) and into pinned memory, with a
    padded leading dimension; the large shape takes the block path, the small one the plain path."""
    import ctypes as C
    from simspread_b200._lib import SS_PREDICT_CLEAN, check
    Xq, Xs, Y = o.synth_dense(nq, ns, nf, nt, seed=9, y_density=0.05, alpha=0.3, weighted=True)
    Y[:, 1 % nt] = 0.0
    ctx, L = ss.Context.default(), ss.lib()
    dq, dx, dy = (ss.DMat.from_host(ctx, a) for a in (Xq, Xs, Y))
    R = ss.DMat(ctx, nq, nt)
    check(L.ss_predict_query(ctx.h, dq.h, dx.h, dy.h, R.h, SS_PREDICT_CLEAN, None))
    want = R.to_host()
    sub = slice(0, min(nq, 64))
    ref = o.predict_blocks_query(Xq[sub], Xs, Y)
    o.clean_blocks(ref, o.degrees_blocks(Xs, Y)[2])
    assert relerr(want[sub], ref) < RTOL
    # pageable destination, ld > rows
    R2 = ss.DMat(ctx, nq, nt)
    out = np.full((nq + 5, nt), 7.0, order="F")
    check(L.ss_predict_query_fetch(ctx.h, dq.h, dx.h, dy.h, R2.h, SS_PREDICT_CLEAN, out.ctypes.data, nq + 5))
    assert np.array_equal(out[:nq], want) and np.all(out[nq:] == 7.0)
    assert np.array_equal(R2.to_host(), want)  # the device copy is complete as well
    # pinned destination
    p = C.c_void_p()
    check(L.ss_host_alloc(nq * nt * 8, C.byref(p)))
    try:
        check(L.ss_predict_query_fetch(ctx.h, dq.h, dx.h, dy.h, R2.h, SS_PREDICT_CLEAN, p, nq))
        got = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(nt, nq)).T
        assert np.array_equal(got, want)
    finally:
        L.ss_host_free(p)
    # the host layer takes this route for the whole query block
    qn, sn = [f"q{i}" for i in range(nq)], [f"s{i}" for i in range(ns)]
    fn, tn = [f"f{i}" for i in range(nf)], [f"t{i}" for i in range(nt)]
    A, B = ss.construct((ss.NamedArray(Y, (sn, tn)), ss.NamedArray(np.zeros((nq, nt)), (qn, tn))),
                        (ss.NamedArray(Xs, (sn, fn)), ss.NamedArray(Xq, (qn, fn))))
    yh = ss.predict((A, B), ss.NamedArray(np.zeros((nq, nt)), (qn, tn)), clean=True, layout="dense")
    assert np.array_equal(yh.array, want)


@pytest.mark.parametrize("nq,ns,nf,nt", [(200, 300, 260, 170), (64, 1000, 129, 515), (33, 70, 64, 128), (10, 65, 1, 1)])
def test_first_product_from_sparse_labels(ss, o, nq, ns, nf, nt):
    """csrc/ss_tsparse.cu: T = (Xs' * (Y ./ ks)) ./ kf from the edge list of the label matrix (what the chain runs for a
    large problem with sparse labels; forced here with SS_T_FORM=sparse) against the dense DMMA form and the oracle:
    ragged tiles, a target without sources, a source without targets, a feature nobody has, NaN / Inf labels (edges
    whose weight spread() sets to 0), and the decline on non-finite feature weights."""
    from simspread_b200._lib import SS_PREDICT_CLEAN, check
    Xq, Xs, Y = o.synth_dense(nq, ns, nf, nt, seed=nq + nt, y_density=0.06, alpha=0.3, weighted=True)
    Y[:, 0] = 0.0
    Y[min(5, ns - 1), :] = 0.0
    Xs[:, nf // 2] = 0.0
    if nt > 3:
        Y[min(7, ns - 1), 2] = np.nan
        Y[min(9, ns - 1), 3] = np.inf
    ctx = ss.Context.default()
    dq, dx, dy = (ss.DMat.from_host(ctx, a) for a in (Xq, Xs, Y))
    res = {}
    for form in ("dense", "sparse"):
        os.environ["SS_T_FORM"] = form
        try:
            R = ss.DMat(ctx, nq, nt)
            l0 = ctx.launch_count()
            check(ss.lib().ss_predict_query(ctx.h, dq.h, dx.h, dy.h, R.h, SS_PREDICT_CLEAN, None))
            res[form] = (R.to_host(), ctx.launch_count() - l0)
        finally:
            os.environ.pop("SS_T_FORM", None)
    assert res["sparse"][1] != res["dense"][1]  # the edge-list form really ran (scan, transpose, fill, product)
    want = o.predict_blocks_query(Xq, Xs, Y)
    o.clean_blocks(want, o.degrees_blocks(Xs, Y)[2])
    for form in ("dense", "sparse"):
        got = res[form][0]
        assert np.array_equal(got == -99, want == -99)
        assert relerr(got, want) < RTOL
    # two runs of the edge-list form are bit-identical (fixed summation order), and independent of the query rows
    os.environ["SS_T_FORM"] = "sparse"
    try:
        R2 = ss.DMat(ctx, nq, nt)
        check(ss.lib().ss_predict_query(ctx.h, dq.h, dx.h, dy.h, R2.h, SS_PREDICT_CLEAN, None))
        assert np.array_equal(R2.to_host(), res["sparse"][0], equal_nan=True)
        # a non-finite feature weight: 0 * Inf = NaN in the dense product -> the edge-list form declines
        Xs2 = Xs.copy()
        Xs2[0, 0] = np.inf
        dx2 = ss.DMat.from_host(ctx, Xs2)
        R3, R4 = ss.DMat(ctx, nq, nt), ss.DMat(ctx, nq, nt)
        check(ss.lib().ss_predict_query(ctx.h, dq.h, dx2.h, dy.h, R3.h, SS_PREDICT_CLEAN, None))
        os.environ["SS_T_FORM"] = "dense"
        check(ss.lib().ss_predict_query(ctx.h, dq.h, dx2.h, dy.h, R4.h, SS_PREDICT_CLEAN, None))
        assert np.array_equal(R3.to_host(), R4.to_host(), equal_nan=True)
    finally:
        os.environ.pop("SS_T_FORM", None)


def test_gemm_k_blocking_switch(ss, o):
    """SS_GEMM_KBLOCK (experiment switch, off by default) runs the second product as accumulating launches over blocks
    of the feature dimension; clean! is applied by the last block only."""
    from simspread_b200._lib import SS_PREDICT_CLEAN, check
    nq, ns, nf, nt = 300, 200, 1100, 260
    Xq, Xs, Y = o.synth_dense(nq, ns, nf, nt, seed=4, y_density=0.05, alpha=0.3, weighted=True)
    Y[:, 3] = 0.0
    want = o.predict_blocks_query(Xq, Xs, Y)
    o.clean_blocks(want, o.degrees_blocks(Xs, Y)[2])
    ctx = ss.Context.default()
    dq, dx, dy = (ss.DMat.from_host(ctx, a) for a in (Xq, Xs, Y))
    R = ss.DMat(ctx, nq, nt)
    os.environ["SS_GEMM_KBLOCK"] = "256"
    try:
        l0 = ctx.launch_count()
        check(ss.lib().ss_predict_query(ctx.h, dq.h, dx.h, dy.h, R.h, SS_PREDICT_CLEAN, None))
        blocked_launches = ctx.launch_count() - l0
    finally:
        os.environ.pop("SS_GEMM_KBLOCK", None)
    l0 = ctx.launch_count()
    R1 = ss.DMat(ctx, nq, nt)
    check(ss.lib().ss_predict_query(ctx.h, dq.h, dx.h, dy.h, R1.h, SS_PREDICT_CLEAN, None))
    assert blocked_launches > ctx.launch_count() - l0
    got = R.to_host()
    assert relerr(got, want) < RTOL
    assert np.array_equal(got == -99, want == -99)


@pytest.mark.parametrize("M,N,K", [(600, 1250, 130), (1650, 1660, 100), (1790, 1900, 70), (2400, 1000, 4000), (2040, 2048, 80),
                                   (1928, 1300, 90), (424, 700, 300)])
@pytest.mark.parametrize("op", ["N", "T"])
def test_gemm_partial_wave_in_row_bands_is_bit_identical(ss, M, N, K, op):
    """The tiles of a partial last wave run as 2 or 4 row bands (csrc/ss_gemm.cu, Unit): 50 tiles -> halves, 169 = 148 +
    21 -> quarters, 210 = 148 + 62 -> halves, 152 = 148 + 4 at a long K, 256 = 148 + 108 -> all tiles in quarters (several
    bands per CTA); 1928 and 424 rows end in a row tile with 8 / 40 valid rows, which gets one / two quarter bands only.  Every entry accumulates the same DMMA sequence,
    so the result must equal the whole-tile launch (SS_GEMM_TAIL_SPLIT=0) bit for bit, epilogues included."""
    from simspread_b200._lib import SS_OP_N, SS_OP_T, check
    rng = np.random.default_rng(M + N + K)
    ctx = ss.Context.default()
    A, B = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    dA = ss.DMat.from_host(ctx, A if op == "N" else A.T)
    dB = ss.DMat.from_host(ctx, B)
    dd = ss.DIVec.from_host(ctx, rng.integers(0, 4, size=M).astype(np.int32))
    df = ss.DIVec.from_host(ctx, rng.integers(0, 2, size=N).astype(np.int32))
    o_ = SS_OP_N if op == "N" else SS_OP_T
    res = {}
    # "0": whole tiles on the whole-tile kernel (what large shapes run); "0b": whole tiles on the band kernel
    for split in ("0", "0b", "2", "4"):
        os.environ["SS_GEMM_TAIL_SPLIT"] = split[0]
        if split == "0b":
            os.environ["SS_GEMM_KERNEL"] = "bands"
        try:
            dC, dE = ss.DMat(ctx, M, N), ss.DMat(ctx, M, N)
            check(ss.lib().ss_gemm_f64(ctx.h, o_, dA.h, dB.h, dC.h, None, None))
            check(ss.lib().ss_gemm_f64(ctx.h, o_, dA.h, dB.h, dE.h, dd.h, df.h))
            res[split] = (dC.to_host(), dE.to_host())
        finally:
            os.environ.pop("SS_GEMM_TAIL_SPLIT", None)
            os.environ.pop("SS_GEMM_KERNEL", None)
    want = A @ B
    bound = np.abs(A) @ np.abs(B)
    assert np.max(np.abs(res["0"][0] - want) / bound) < 1e-13
    for split in ("0b", "2", "4"):
        assert np.array_equal(res[split][0], res["0"][0]) and np.array_equal(res[split][1], res["0"][1])


def test_empty_and_degenerate_inputs(ss, o):
    from simspread_b200._lib import check
    ctx = ss.Context.default()
    # no edges at all: every degree is 0 -> W = 0 -> scores 0, every column flagged by clean!
    Xq, Xs, Y = np.ones((4, 3)), np.zeros((5, 3)), np.zeros((5, 2))
    dq, dx, dy, R = (ss.DMat.from_host(ctx, a) for a in (Xq, Xs, Y, np.ones((4, 2))))
    check(ss.lib().ss_predict_query(ctx.h, dq.h, dx.h, dy.h, R.h, 0, None))
    assert np.array_equal(R.to_host(), np.zeros((4, 2)))
    check(ss.lib().ss_predict_query(ctx.h, dq.h, dx.h, dy.h, R.h, 1, None))
    assert np.array_equal(R.to_host(), np.full((4, 2), -99.0))
    # zero queries / zero-size matrices are accepted
    e = ss.DMat(ctx, 0, 3)
    R0 = ss.DMat(ctx, 0, 2)
    check(ss.lib().ss_predict_query(ctx.h, e.h, dx.h, dy.h, R0.h, 0, None))
    assert ss.cutoff(np.zeros((0, 4)), 0.5).shape == (0, 4)
    # shape errors carry the reference's assertion text
    bad = ss.DMat(ctx, 4, 2)
    st = ss.lib().ss_predict_query(ctx.h, bad.h, dx.h, dy.h, R.h, 0, None)
    assert st != 0 and b"Number of features" in ss.lib().ss_last_error()


# ---------------------------------------------------------------------------------------------
# ranking / metrics
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("rows,cols,L", [(1, 30, 20), (45, 664, 20), (200, 2000, 5), (130, 77, 64)])
def test_topl_order_bit_exact(ss, o, rows, cols, L):
    from simspread_b200._lib import check
    rng = np.random.default_rng(rows + cols + L)
    R = np.round(rng.random((rows, cols)), 2)  # many ties
    R[rng.random((rows, cols)) < 0.3] = 0.0
    R[0, :min(cols, 5)] = [np.nan, -0.0, 0.0, -99.0, np.inf][:min(cols, 5)]
    ctx = ss.Context.default()
    d = ss.DMat.from_host(ctx, R)
    idx, val = ss.DIVec(ctx, L * rows), ss.DMat(ctx, L, rows)
    check(ss.lib().ss_topl_rows(ctx.h, d.h, L, idx.h, val.h))
    got = idx.to_host().reshape(rows, L)
    want = np.stack([o.sortperm_rev(R[r])[:L] for r in range(rows)])
    assert np.array_equal(got, want)
    gv = val.to_host().T
    wv = np.take_along_axis(R, want, axis=1)
    assert np.array_equal(gv, wv, equal_nan=True)


def test_atl_against_oracle(ss, o):
    rng = np.random.default_rng(11)
    rows, cols, L = 60, 300, 20
    R = np.round(rng.random((rows, cols)), 2)
    Y = (rng.random((rows, cols)) < 0.05).astype(float)
    y, s = Y.ravel(order="C"), R.ravel(order="C")
    grp = np.repeat(np.arange(rows), cols)
    assert ss.recallatL(y, s, grp, L) == pytest.approx(o.recallatL_grouped(y, s, grp, L), rel=1e-12, nan_ok=True)
    assert ss.precisionatL(y, s, grp, L) == pytest.approx(o.precisionatL_grouped(y, s, grp, L), rel=1e-12)
    Y[3, :] = 0  # a group without positives: NaN propagates through the mean
    y = Y.ravel(order="C")
    assert math.isnan(ss.recallatL(y, s, grp, L)) and math.isnan(o.recallatL_grouped(y, s, grp, L))
    # ragged groups
    grp2 = np.concatenate([np.zeros(40), np.ones(100), np.full(rows * cols - 140, 2)])
    assert ss.precisionatL(y, s, grp2, L) == pytest.approx(o.precisionatL_grouped(y, s, grp2, L), rel=1e-12)
    assert ss.recallatL(y[:100], s[:100], 7) == pytest.approx(o.recallatL(y[:100], s[:100], 7), rel=1e-12, nan_ok=True)


@pytest.mark.parametrize("M,ties", [(2, False), (1000, True), (4096, False), (300000, True), (1 << 20, False)])
def test_auroc_auprc_against_oracle(ss, o, M, ties):
    rng = np.random.default_rng(M)
    y = rng.random(M) < 0.03
    y[0] = True
    y[1] = False
    s = rng.random(M) * (y * 0.3 + 0.7)
    if ties:
        s = np.round(s, 3)
        s[rng.random(M) < 0.5] = 0.0
    a, p = ss.AuROC(y, s), ss.AuPRC(y, s)
    assert a == pytest.approx(o.AuROC(y, s), rel=1e-12)
    assert p == pytest.approx(o.AuPRC(y, s), rel=1e-12)


def test_auroc_quirks_and_edges(ss, o):
    # App. A item 16: no anchors -> not the textbook values
    assert ss.AuROC([1, 0, 1, 0, 0], [.9, .9, .7, .1, .1]) == pytest.approx(2 / 3, rel=1e-12)
    assert ss.AuROC([1, 0, 1], [.5, .5, .5]) == 0.0
    assert ss.AuPRC([1, 0, 1, 0], [.9, .8, .7, .1]) == pytest.approx(0.2916666666666667, rel=1e-12)
    assert math.isnan(ss.AuROC([0, 0, 0], [.1, .2, .3]))  # no positives: 0/0
    s = np.array([-99.0, 0.0, 0.5, -99.0, 0.25, 0.0])
    y = np.array([0, 1, 1, 0, 0, 0])
    assert ss.AuROC(y, s) == pytest.approx(o.AuROC(y, s), rel=1e-12)
    with pytest.raises(AssertionError, match="The number of scores must be equal"):
        ss.AuROC([1, 0], [0.5])
    assert ss.validity_ratio(s) == o.validity_ratio(s)


def test_large_gemm_properties(ss):
    """Full-width tiles, many k-slabs, multi-wave persistent schedule: integer-valued operands make
    the result independent of summation order, so it must equal the CPU product exactly."""
    from simspread_b200._lib import SS_OP_N, check
    rng = np.random.default_rng(99)
    M, N, K = 2500, 3000, 2100
    A = rng.integers(0, 3, size=(M, K)).astype(float)
    B = rng.integers(-2, 3, size=(K, N)).astype(float)
    ctx = ss.Context.default()
    dA, dB, dC = ss.DMat.from_host(ctx, A), ss.DMat.from_host(ctx, B), ss.DMat(ctx, M, N)
    check(ss.lib().ss_gemm_f64(ctx.h, SS_OP_N, dA.h, dB.h, dC.h, None, None))
    assert np.array_equal(dC.to_host(), A @ B)


def test_sharded_c_abi_world1_matches_single_call(ss, o):
    """The N > 1 product path (ss_comm_* / ss_sharded_* / ss_predict_query_sharded behind the C ABI) with a
    one-rank communicator must reproduce ss_predict_query bit for bit; the small collectives degrade to copies."""
    from simspread_b200._lib import SS_PREDICT_CLEAN, check
    from simspread_b200.sharded import Comm, ShardedQuery
    nq, ns, nf, nt = 200, 150, 130, 90
    Xq, Xs, Y = o.synth_dense(nq, ns, nf, nt, seed=9, y_density=0.05, alpha=0.2, weighted=True)
    Y[:, 3] = 0.0
    ctx = ss.Context.default()
    L = ss.lib()
    comm = Comm(ctx, 0, 1)
    plan = ShardedQuery(comm, ns, nf, nt)
    assert plan.nt_blk == nt and not plan.fused
    dq, dx, dy = (ss.DMat.from_host(ctx, a) for a in (Xq, Xs, Y))
    R1, R2 = ss.DMat(ctx, nq, nt), ss.DMat(ctx, nq, nt)
    plan.predict(dq, dx, dy, R1, clean=True)
    check(L.ss_predict_query(ctx.h, dq.h, dx.h, dy.h, R2.h, SS_PREDICT_CLEAN, None))
    got = R1.to_host()
    assert np.array_equal(got, R2.to_host())
    want = o.predict_blocks_query(Xq, Xs, Y)
    o.clean_blocks(want, o.degrees_blocks(Xs, Y)[2])
    assert relerr(got, want) < RTOL
    # a rank without query rows still takes part in the front; T / kt views feed the streaming product
    plan.predict(None, dx, dy, None)
    hT, hkt = plan.views()
    Rh = np.zeros((nq, nt), order="F")
    Xqf = np.asfortranarray(Xq)
    check(L.ss_stream_product_host(ctx.h, Xqf.ctypes.data, nq, nq, hT, hkt, Rh.ctypes.data, nq))
    assert np.array_equal(Rh, got)
    # the helper collectives on one rank
    v = ss.DIVec.from_host(ctx, np.arange(7, dtype=np.int32))
    check(L.ss_comm_allreduce_i32(comm.h, v.h))
    w = ss.DIVec(ctx, 7)
    check(L.ss_comm_allgather_i32(comm.h, v.h, w.h))
    assert np.array_equal(w.to_host(), np.arange(7))
    assert comm.allreduce_host([1.5, -2.0], "max") == [1.5, -2.0]
    check(L.ss_comm_allgather_cols(comm.h, dx.h))
    comm.barrier()
    plan.close()
    comm.close()


def _sharded_rank_main(rank, world, path, nq, ns, nf, nt, out_dir):
    """One rank of the 2-GPU C-ABI test (spawned process): file rendezvous, sharded predict, result to disk."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import simspread_b200 as m
    from oracle import simspread_oracle as oo
    from simspread_b200.sharded import Comm, ShardedQuery
    ctx = m.Context(rank)
    Xq, Xs, Y = oo.synth_dense(nq, ns, nf, nt, seed=13, y_density=0.05, alpha=0.2, weighted=True)
    Y[:, 5] = 0.0
    comm = Comm(ctx, rank, world, path=path)
    plan = ShardedQuery(comm, ns, nf, nt)
    blk = plan.nt_blk
    nq_blk = -(-nq // world)
    q0, q1 = min(nq, rank * nq_blk), min(nq, (rank + 1) * nq_blk)
    Yb = np.zeros((ns, blk))
    c0, c1 = rank * blk, min(nt, (rank + 1) * blk)
    Yb[:, :max(0, c1 - c0)] = Y[:, c0:c1]
    dx, dy = m.DMat.from_host(ctx, Xs), m.DMat.from_host(ctx, Yb)
    if q1 > q0:
        dq, R = m.DMat.from_host(ctx, Xq[q0:q1]), m.DMat(ctx, q1 - q0, nt)
        for _ in range(2):  # twice: the second front overwrites T while the peers may still hold the first
            plan.predict(dq, dx, dy, R, clean=True)
        res = R.to_host()
    else:
        for _ in range(2):
            plan.predict(None, dx, dy, None)
        res = np.zeros((0, nt))
    np.save(os.path.join(out_dir, f"r{rank}.npy"), res)
    with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as f:
        f.write(f"{int(plan.fused)} {comm.nccl_version()}")
    plan.close()
    comm.close()


@pytest.mark.parametrize("tform", ["auto", "sparse"])
@pytest.mark.parametrize("fused", ["1", "0"])
def test_sharded_c_abi_two_gpus(ss, o, tmp_path, fused, tform):
    """Two processes, two GPUs, nothing but the C ABI between them (NCCL dlopen()ed by the library, unique id through
    a file, T tiles stored into the peer from the GEMM epilogue or all-gathered by NCCL): the row slabs must equal
    the single-GPU result bit for bit (same kernels, same order of additions) and the oracle within 1e-12."""
    import multiprocessing as mp
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    nq, ns, nf, nt = 301, 150, 130, 91  # ragged: the last rank owns fewer rows / target columns
    os.environ["SS_FUSED_ALLGATHER"] = fused
    if tform == "sparse":  # the edge-list form of the first product (csrc/ss_tsparse.cu), T tiles stored into the peer
        os.environ["SS_T_FORM"] = "sparse"
    try:
        mpc = mp.get_context("spawn")
        path = str(tmp_path / "nccl_id")
        procs = [mpc.Process(target=_sharded_rank_main, args=(r, 2, path, nq, ns, nf, nt, str(tmp_path))) for r in range(2)]
        for p_ in procs:
            p_.start()
        for p_ in procs:
            p_.join(300)
            assert p_.exitcode == 0
    finally:
        os.environ.pop("SS_FUSED_ALLGATHER", None)
        os.environ.pop("SS_T_FORM", None)
    got = np.concatenate([np.load(tmp_path / f"r{r}.npy") for r in range(2)])
    flags = [open(tmp_path / f"r{r}.txt").read().split() for r in range(2)]
    assert all(f[0] == fused for f in flags) or fused == "1"  # fused falls back to NCCL without peer access
    Xq, Xs, Y = o.synth_dense(nq, ns, nf, nt, seed=13, y_density=0.05, alpha=0.2, weighted=True)
    Y[:, 5] = 0.0
    from simspread_b200._lib import SS_PREDICT_CLEAN, check
    ctx = ss.Context.default()
    dq, dx, dy, R = (ss.DMat.from_host(ctx, a) for a in (Xq, Xs, Y, np.zeros((nq, nt))))
    if tform == "sparse":
        os.environ["SS_T_FORM"] = "sparse"
    try:
        check(ss.lib().ss_predict_query(ctx.h, dq.h, dx.h, dy.h, R.h, SS_PREDICT_CLEAN, None))
    finally:
        os.environ.pop("SS_T_FORM", None)
    assert np.array_equal(got, R.to_host())
    want = o.predict_blocks_query(Xq, Xs, Y)
    o.clean_blocks(want, o.degrees_blocks(Xs, Y)[2])
    assert relerr(got, want) < RTOL


@pytest.mark.parametrize("N,Nt,k_,weighted", [(445, 664, 10, False), (130, 70, 4, True), (67, 129, 67, True)])
def test_folds_in_three_launches_match_the_per_fold_chain(ss, o, monkeypatch, N, Nt, k_, weighted):
    """Small folds run as ONE grid over (fold, tile) -- degrees, T and R of every fold in three launches, operands read
    through the fold index lists (ss_folds.cu).  Same predictions as the per-fold chain (gather, degrees, spread, two
    DMMA GEMMs per fold) within the FP64 tolerance, same -99 flags, and both match the oracle loop."""
    rng = np.random.default_rng(N + Nt)
    S = np.round(rng.random((N, N)), 4)
    Yf = (rng.random((N, Nt)) < 0.03).astype(float)
    Yf[:, 2] = 0.0
    names = [f"D{i:04d}" for i in range(N)]
    tn = [f"T{j:04d}" for j in range(Nt)]
    DD, DT = ss.NamedArray(S, (names, names)), ss.NamedArray(Yf, (names, tn))
    ctx = ss.Context.default()
    l0 = ctx.launch_count()
    a = ss.cross_validate(DT, DD, 0.45, weighted=weighted, k_=k_, seed=3)
    launches_batched = ctx.launch_count() - l0
    monkeypatch.setenv("SS_FOLDS_SERIAL", "1")
    l0 = ctx.launch_count()
    b = ss.cross_validate(DT, DD, 0.45, weighted=weighted, k_=k_, seed=3)
    launches_serial = ctx.launch_count() - l0
    monkeypatch.delenv("SS_FOLDS_SERIAL")
    assert a["folds"] == b["folds"]
    assert relerr(a["yhat"].array, b["yhat"].array) < RTOL
    assert np.array_equal(a["yhat"].array == -99, b["yhat"].array == -99)
    assert launches_serial - launches_batched >= 5 * k_ - 3   # 7 launches per fold became 3 per CV
    Xo, xr, xc = o.featurize(S, names, names, 0.45, weighted)
    row = 0
    for q in a["folds"]:
        qi = [names.index(x) for x in q]
        si = [i for i in range(N) if names[i] not in set(q)]
        want = o.predict_blocks_query(Xo[np.ix_(qi, si)], Xo[np.ix_(si, si)], Yf[si])
        o.clean_blocks(want, o.degrees_blocks(Xo[np.ix_(si, si)], Yf[si])[2])
        assert relerr(a["yhat"].array[row:row + len(q)], want) < RTOL
        row += len(q)


def test_cross_validate_enzyme_shape_against_oracle_loop(ss, o):
    """BASELINE config 2: Enzyme-shaped (445 x 664), binary alpha-cutoff features, 10-fold CV."""
    S, Yfull = _enzyme_like(o, seed=20242)
    N, Nt = Yfull.shape
    names = [f"D{i:04d}" for i in range(N)]
    tnames = [f"T{j:04d}" for j in range(Nt)]
    DD, DT = ss.NamedArray(S, (names, names)), ss.NamedArray(Yfull, (names, tnames))
    res = ss.cross_validate(DT, DD, 0.35, weighted=False, k_=10, seed=1, L=20)
    folds = res["folds"]
    assert len(folds) == 10 and sorted(sum(folds, [])) == sorted(names)
    Xo, xr, xc = o.featurize(S, names, names, 0.35, False)
    want, order = [], []
    for q in folds:
        Ao, Bo, nn = o.construct_queries(Yfull, (names, tnames), Xo, (xr, xc), q)
        w = o.predict_dense(Ao, Bo, nn, q, tnames)
        o.clean(w, Ao, nn, tnames)
        want.append(w)
        order += q
    want = np.vstack(want)
    assert res["yhat"].names(1) == order
    assert relerr(res["yhat"].array, want) < RTOL
    assert np.array_equal(res["yhat"].array == -99, want == -99)
    ytrue = Yfull[[names.index(x) for x in order]]
    assert np.array_equal(res["y"].array, ytrue)
    # metrics on the reference-path scores (ties resolve identically only for identical scores)
    assert res["AuROC"] == pytest.approx(o.AuROC(ytrue.ravel() > 0, res["yhat"].array.ravel()), rel=1e-12)
    assert res["AuPRC"] == pytest.approx(o.AuPRC(ytrue.ravel() > 0, res["yhat"].array.ravel()), rel=1e-12)
    grp = np.repeat(np.arange(N), Nt)
    sc = res["yhat"].array.ravel()
    assert res["precisionatL"] == pytest.approx(o.precisionatL_grouped(ytrue.ravel(), sc, grp, 20), rel=1e-12)
    r = o.recallatL_grouped(ytrue.ravel(), sc, grp, 20)
    assert (math.isnan(r) and math.isnan(res["recallatL"])) or res["recallatL"] == pytest.approx(r, rel=1e-12)


def test_cross_validate_folds_sharded_over_ranks(ss, o):
    """SURVEY 8e, outer level: CV folds are independent units; rank r takes folds[r::world], the parts are
    gathered on the host and merged -- same predictions and metrics as the single-GPU call, bit for bit."""
    S, Yfull = _enzyme_like(o, seed=20243)
    N, Nt = Yfull.shape
    names = [f"D{i:04d}" for i in range(N)]
    tnames = [f"T{j:04d}" for j in range(Nt)]
    DD, DT = ss.NamedArray(S, (names, names)), ss.NamedArray(Yfull, (names, tnames))
    one = ss.cross_validate(DT, DD, 0.35, weighted=False, k_=10, seed=3, L=20)
    world = 3
    parts = [ss.cross_validate(DT, DD, 0.35, weighted=False, k_=10, seed=3, L=20, rank=r, world=world) for r in range(world)]
    assert [p["fold_ids"] for p in parts] == [[0, 3, 6, 9], [1, 4, 7], [2, 5, 8]]
    assert all("AuROC" not in p for p in parts)
    merged = ss.merge_cross_validation(parts[::-1], L=20)  # gather order must not matter
    assert merged["folds"] == one["folds"] and merged["yhat"].names(1) == one["yhat"].names(1)
    assert np.array_equal(merged["yhat"].array, one["yhat"].array)
    assert np.array_equal(merged["y"].array, one["y"].array)
    for key in ("AuROC", "AuPRC", "precisionatL"):
        assert merged[key] == one[key]
    assert merged["recallatL"] == one["recallatL"] or (math.isnan(merged["recallatL"]) and math.isnan(one["recallatL"]))


def test_alpha_sweep_against_oracle(ss, o):
    """BASELINE config 3 (scaled): weighted SimSpread over alpha in 0..1, dense end to empty end."""
    rng = np.random.default_rng(33)
    nq, ns, nt = 120, 260, 150
    N = nq + ns
    names = [f"n{i}" for i in range(N)]
    tn = [f"t{j}" for j in range(nt)]
    S = np.round(rng.random((N, N)), 6)
    Yl = (rng.random((N, nt)) < 0.02).astype(float)
    DD, DT = ss.NamedArray(S, (names, names)), ss.NamedArray(Yl, (names, tn))
    queries = names[:nq]
    alphas = [0.0, 0.25, 0.5, 0.9, 1.0]
    got = ss.alpha_sweep(DT, DD, queries, alphas, weighted=True, L=20)
    both = ss.alpha_sweep(DT, DD, queries, alphas, weighted=True, L=20, rank=1, world=2)
    assert [g["alpha"] for g in both] == alphas[1::2]
    for g in got:
        Xo, xr, xc = o.featurize(S, names, names, g["alpha"], True)
        Ao, Bo, nn = o.construct_queries(Yl, (names, tn), Xo, (xr, xc), queries)
        w = o.predict_dense(Ao, Bo, nn, queries, tn)
        o.clean(w, Ao, nn, tn)
        yq = Yl[:nq]
        # score ties at exactly-equal sums can differ in the last bit between the two summation
        # orders, so the metric is compared on the oracle's own scores with a small tolerance
        assert g["AuROC"] == pytest.approx(o.AuROC(yq.ravel() > 0, w.ravel()), rel=1e-9, nan_ok=True)
        assert g["AuPRC"] == pytest.approx(o.AuPRC(yq.ravel() > 0, w.ravel()), rel=1e-9, nan_ok=True)
        assert g["validity_ratio"] == o.validity_ratio(w)
    assert got[-1]["validity_ratio"] <= got[0]["validity_ratio"]


@pytest.mark.parametrize("rows,cols", [(1, 1), (300, 7), (445, 445), (1000, 333)])
@pytest.mark.parametrize("weighted", [False, True])
def test_featurize_csc_bit_exact(ss, o, rows, cols, weighted):
    rng = np.random.default_rng(rows + 17 * cols)
    S = np.round(rng.random((rows, cols)), 3)
    S.flat[rng.integers(0, S.size, size=max(1, S.size // 50))] = np.nan
    S.flat[rng.integers(0, S.size, size=max(1, S.size // 50))] = 0.0
    ctx = ss.Context.default()
    d = ss.DMat.from_host(ctx, S)
    for alpha in (0.35, 0.0, 0.95, 1.01):
        want = o.cutoff(S, alpha, weighted)
        c = ss.DCsr.from_dense(ctx, d, alpha, weighted, by_columns=True)
        rp, ci, va = c.to_host()
        wc, wr = np.nonzero(want.T != 0)  # CSR of S': row = column of S, ascending row index
        assert (c.rows, c.cols, c.nnz) == (cols, rows, len(wr))
        assert np.array_equal(rp, np.concatenate(([0], np.cumsum(np.bincount(wc, minlength=cols)))).astype(np.int32))
        assert np.array_equal(ci, wr.astype(np.int32))
        if weighted:
            assert np.array_equal(va, want[wr, wc])
        else:
            assert va is None


@pytest.mark.parametrize("weighted,alpha", [(True, 0.9), (False, 0.97), (True, 0.5), (True, 1.01)])
def test_predict_sparse_chain_against_oracle_and_dense(ss, o, weighted, alpha):
    from simspread_b200._lib import SS_PREDICT_CLEAN, check
    nq, ns, nf, nt = 333, 410, 410, 517
    rng = np.random.default_rng(int(alpha * 100))
    Xq = o.cutoff(np.round(rng.random((nq, nf)), 6), alpha, weighted)
    Xs = o.cutoff(np.round(rng.random((ns, nf)), 6), alpha, weighted)
    Y = (rng.random((ns, nt)) < 0.03).astype(float)
    Y[:, 5] = 0.0
    want = o.predict_blocks_query(Xq, Xs, Y)
    o.clean_blocks(want, o.degrees_blocks(Xs, Y)[2])
    ctx = ss.Context.default()
    dq, dx, dy = (ss.DMat.from_host(ctx, a) for a in (Xq, Xs, Y))
    cq = ss.DCsr.from_dense(ctx, dq, float("-inf"), True)
    cs = ss.DCsr.from_dense(ctx, dx, float("-inf"), True, by_columns=True)
    assert cq.nnz == np.count_nonzero(Xq) and cs.nnz == np.count_nonzero(Xs)
    R = ss.DMat.from_host(ctx, np.full((nq, nt), 3.0))
    kt = ss.DIVec(ctx, nt)
    check(ss.lib().ss_predict_query_csr(ctx.h, cq.h, cs.h, dy.h, R.h, SS_PREDICT_CLEAN, kt.h))
    got = R.to_host()
    assert relerr(got, want) < RTOL and np.array_equal(got == -99, want == -99)
    assert np.array_equal(kt.to_host(), o.degrees_blocks(Xs, Y)[2])
    Rd = ss.DMat(ctx, nq, nt)
    check(ss.lib().ss_predict_query(ctx.h, dq.h, dx.h, dy.h, Rd.h, SS_PREDICT_CLEAN, None))
    assert relerr(Rd.to_host(), want) < RTOL


def test_predict_layout_auto_switch(ss, o):
    rng = np.random.default_rng(4)
    N, nt = 400, 90
    names = [f"n{i}" for i in range(N)]
    tn = [f"t{j}" for j in range(nt)]
    S = np.round(rng.random((N, N)), 6)
    Yl = (rng.random((N, nt)) < 0.05).astype(float)
    DT = ss.NamedArray(Yl, (names, tn))
    q = names[:80]
    for alpha, expect in ((0.3, "dense"), (0.985, "sparse")):
        X = ss.featurize(ss.NamedArray(S, (names, names)), alpha, True)
        A, B = ss.construct(DT, X, q)
        auto = ss.predict((A, B), DT[q, tn])
        assert A.last_layout == expect
        dense = ss.predict((A, B), DT[q, tn], layout="dense")
        sparse = ss.predict((A, B), DT[q, tn], layout="sparse")
        Ao, Bo, nn = o.construct_queries(Yl, (names, tn), X.array, (names, X.names(2)), q)
        want = o.predict_dense(Ao, Bo, nn, q, tn)
        for got in (auto, dense, sparse):
            assert relerr(got.array, want) < RTOL


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (129, 257, 33), (300, 520, 260), (45, 664, 400), (1000, 1500, 2000)])
@pytest.mark.parametrize("op", ["N", "T"])
def test_gemm_tf32_tcgen05(ss, M, N, K, op):
    """Opt-in tcgen05 path: exact on operands that are exactly representable in TF32 (a descriptor or
    pipeline bug cannot hide behind rounding), and within the stated TF32 bound on real data."""
    from simspread_b200._lib import SS_OP_N, SS_OP_T, SS_PRECISION_TF32, check
    rng = np.random.default_rng(M + 3 * N + 5 * K)
    ctx = ss.Context.default()
    o_ = SS_OP_N if op == "N" else SS_OP_T
    A = rng.integers(-3, 4, size=(M, K)).astype(float)
    B = rng.integers(-3, 4, size=(K, N)).astype(float)
    div = rng.integers(0, 3, size=M).astype(np.int32)
    flag = rng.integers(0, 2, size=N).astype(np.int32)
    dA, dB, dC = ss.DMat.from_host(ctx, A if op == "N" else A.T), ss.DMat.from_host(ctx, B), ss.DMat(ctx, M, N)
    check(ss.lib().ss_gemm_lowp(ctx.h, o_, dA.h, dB.h, dC.h, None, None, SS_PRECISION_TF32))
    assert np.array_equal(dC.to_host(), A @ B)
    dd, df = ss.DIVec.from_host(ctx, div), ss.DIVec.from_host(ctx, flag)
    check(ss.lib().ss_gemm_lowp(ctx.h, o_, dA.h, dB.h, dC.h, dd.h, df.h, SS_PRECISION_TF32))
    with np.errstate(divide="ignore", invalid="ignore"):
        ref = np.where(div[:, None] == 0, 0.0, (A @ B) / div[:, None].astype(float))
    ref[:, flag == 0] = -99.0
    assert np.allclose(dC.to_host(), ref, rtol=1e-6, atol=0)
    A, B = rng.random((M, K)), rng.random((K, N))
    dA, dB = ss.DMat.from_host(ctx, A if op == "N" else A.T), ss.DMat.from_host(ctx, B)
    check(ss.lib().ss_gemm_lowp(ctx.h, o_, dA.h, dB.h, dC.h, None, None, SS_PRECISION_TF32))
    assert np.max(np.abs(dC.to_host() - A @ B) / (A @ B)) < 2e-3  # TF32: 10-bit mantissa operands


def test_predict_tf32_mode_against_oracle(ss, o):
    S, Yfull = _enzyme_like(o, seed=5)
    N, Nt = Yfull.shape
    names = [f"D{i:04d}" for i in range(N)]
    tn = [f"T{j:04d}" for j in range(Nt)]
    DT = ss.NamedArray(Yfull, (names, tn))
    X = ss.featurize(ss.NamedArray(S, (names, names)), 0.2, True)
    q = names[::9]
    A, B = ss.construct(DT, X, q)
    got = ss.predict((A, B), DT[q, tn], precision="tf32", clean=True)
    exact = ss.predict((A, B), DT[q, tn], clean=True)
    nz = (exact.array != 0) & (exact.array != -99)
    assert np.array_equal(got.array == -99, exact.array == -99)
    assert np.max(np.abs(got.array[nz] - exact.array[nz]) / exact.array[nz]) < 2e-3  # stated TF32 bound
    assert np.array_equal(got.array[exact.array == 0], exact.array[exact.array == 0])
    with pytest.raises(ValueError):
        ss.predict((A, B), DT[q, tn], precision="f16")


def test_bedroc_and_threshold_sweeps_against_oracle(ss, o):
    rng = np.random.default_rng(8)
    M = 20000
    y = rng.random(M) < 0.05
    s = np.round(rng.random(M) * (y * 0.4 + 0.6), 3)  # ties
    s[rng.random(M) < 0.3] = 0.0
    assert ss.BEDROC(y, s) == pytest.approx(o.BEDROC(y, s), rel=1e-12)
    assert ss.BEDROC(y, s, rev=False, alpha=5.0) == pytest.approx(o.BEDROC(y, s, rev=False, alpha=5.0), rel=1e-12)
    pairs = [(ss.f1score, o.f1score), (ss.mcc, o.mcc), (ss.accuracy, o.accuracy),
             (ss.balancedaccuracy, o.balancedaccuracy), (ss.recall, o.recall), (ss.precision, o.precision)]
    for mine, ref in pairs:
        assert ss.maxperformance(y, s, mine) == pytest.approx(o.maxperformance(y, s, ref), rel=1e-12, nan_ok=True)
        assert ss.meanperformance(y, s, mine) == pytest.approx(o.meanperformance(y, s, ref), rel=1e-11, nan_ok=True)
        m, sd = ss.meanstdperformance(y, s, mine)
        mo, sdo = o.meanstdperformance(y, s, ref)
        assert m == pytest.approx(mo, rel=1e-11, nan_ok=True) and sd == pytest.approx(sdo, rel=1e-9, nan_ok=True)
    # a metric that is NaN at some threshold poisons max / mean, as Julia's maximum / mean do
    yn = np.zeros(50, bool)
    assert math.isnan(ss.maxperformance(yn, rng.random(50), ss.recall))


# (1700, 3100, 300): 14 x 13 = 182 tiles > 148 SMs -- the persistent tile loop and the producers' lockstep
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (129, 257, 133), (300, 520, 2600), (64, 200, 20000), (1700, 3100, 300)])
@pytest.mark.parametrize("op", ["N", "T"])
def test_gemm_f64_from_int8_slices(ss, M, N, K, op):
    """Opt-in FP64-grade mode on the INT8 tensor pipe: exact integer slice products, FP64
    recombination.  Exact on integer operands; within the FP64 score tolerance on real data."""
    from simspread_b200._lib import SS_OP_N, SS_OP_T, SS_PRECISION_F64_INT8, check
    rng = np.random.default_rng(M + 3 * N + 7 * K)
    ctx = ss.Context.default()
    o_ = SS_OP_N if op == "N" else SS_OP_T
    A = rng.integers(0, 250, size=(M, K)).astype(float)
    B = rng.integers(0, 250, size=(K, N)).astype(float)
    if M > 1:
        A[0, :] = 0.0  # an all-zero row (scale 2^0, all digits 0)
    dA, dB, dC = ss.DMat.from_host(ctx, A if op == "N" else A.T), ss.DMat.from_host(ctx, B), ss.DMat(ctx, M, N)
    check(ss.lib().ss_gemm_lowp(ctx.h, o_, dA.h, dB.h, dC.h, None, None, SS_PRECISION_F64_INT8))
    assert np.array_equal(dC.to_host(), A @ B)
    # rows / columns of very different magnitude: the per-row power-of-two scaling keeps them exact
    A = rng.random((M, K)) * np.exp2(rng.integers(-30, 30, size=(M, 1)))
    B = rng.random((K, N)) * np.exp2(rng.integers(-30, 30, size=(1, N)))
    div = rng.integers(0, 3, size=M).astype(np.int32)
    flag = rng.integers(0, 2, size=N).astype(np.int32)
    dA, dB = ss.DMat.from_host(ctx, A if op == "N" else A.T), ss.DMat.from_host(ctx, B)
    check(ss.lib().ss_gemm_lowp(ctx.h, o_, dA.h, dB.h, dC.h, None, None, SS_PRECISION_F64_INT8))
    want = (A.astype(np.longdouble) @ B.astype(np.longdouble)).astype(float)
    assert np.max(np.abs(dC.to_host() - want) / want) < RTOL
    dd, df = ss.DIVec.from_host(ctx, div), ss.DIVec.from_host(ctx, flag)
    check(ss.lib().ss_gemm_lowp(ctx.h, o_, dA.h, dB.h, dC.h, dd.h, df.h, SS_PRECISION_F64_INT8))
    with np.errstate(divide="ignore", invalid="ignore"):
        ref = np.where(div[:, None] == 0, 0.0, want / div[:, None].astype(float))
    ref[:, flag == 0] = -99.0
    got = dC.to_host()
    assert np.array_equal(got == -99, ref == -99) and np.array_equal(got == 0, ref == 0)
    nz = (ref != 0) & (ref != -99)
    assert not nz.any() or np.max(np.abs(got[nz] - ref[nz]) / ref[nz]) < RTOL
    # negative / non-finite operands are refused (the slicing is defined for non-negative graphs)
    A[M // 2, K // 2] = -1.0
    dA = ss.DMat.from_host(ctx, A if op == "N" else A.T)
    st = ss.lib().ss_gemm_lowp(ctx.h, o_, dA.h, dB.h, dC.h, None, None, SS_PRECISION_F64_INT8)
    assert st == 6 and b"non-negative" in ss.lib().ss_last_error()


def test_int8_mode_certificate_and_fp64_fallback(ss):
    """precision="f64_int8" certifies every entry a posteriori (error bound <= 4e-13 * entry, or an exact zero with no
    operand entry truncated to zero).  Products that cannot be certified are re-run on the FP64 DMMA path, so the
    result is within the FP64 tolerance either way; ss_ctx_int8_stats tells which path produced it."""
    from simspread_b200._lib import SS_OP_N, SS_PRECISION_F64_INT8, check
    rng = np.random.default_rng(99)
    ctx = ss.Context.default()
    M, N, K = 200, 300, 4000

    def run(A, B):
        before = ctx.int8_stats()
        dA, dB, dC = ss.DMat.from_host(ctx, A), ss.DMat.from_host(ctx, B), ss.DMat(ctx, A.shape[0], B.shape[1])
        check(ss.lib().ss_gemm_lowp(ctx.h, SS_OP_N, dA.h, dB.h, dC.h, None, None, SS_PRECISION_F64_INT8))
        after = ctx.int8_stats()
        assert after[0] == before[0] + 1
        want = (A.astype(np.longdouble) @ B.astype(np.longdouble)).astype(float)
        got = dC.to_host()
        assert np.array_equal(got == 0, want == 0)
        nz = want != 0
        assert np.max(np.abs(got[nz] - want[nz]) / want[nz]) < RTOL
        return after[1] - before[1], after[2]

    # (1) dense similarity-like operands: certified, no fallback; 6 x 6 slices with i + j <= 7 -> 21 pairs
    A, B = rng.random((M, K)), rng.random((K, N))
    assert run(A, B) == (0, 0) and ctx.int8_last_pairs() == 21
    # (1b) a 0/1 operand is a single 8-bit plane: planes of zeros are skipped -> 6 pairs, same certified result
    Bbin = (rng.random((K, N)) < 0.05).astype(float)
    assert run(A, Bbin) == (0, 0) and ctx.int8_last_pairs() == 6
    assert run((A > 0.5).astype(float), Bbin) == (0, 0) and ctx.int8_last_pairs() == 1
    # (2) block structure: exact zeros (disjoint supports) are certified as zeros
    A2, B2 = A.copy(), B.copy()
    A2[:, K // 2:] = 0.0
    B2[:K // 2, :N // 2] = 0.0
    assert run(A2, B2) == (0, 0)
    # (3) one huge entry per row next to tiny ones: the normwise bound says nothing about the small entries of the
    #     product -> the certificate fails -> FP64 fallback, result still within tolerance
    A3 = rng.random((M, K)) * 1e-9
    A3[:, 0] = 1.0
    B3 = rng.random((K, N))
    B3[0, :] = 0.0
    fell_back, failed = run(A3, B3)
    assert fell_back == 1 and failed > 0
    # (4) an entry far below 2^-48 of its row maximum is truncated to 0: exact zeros can no longer be certified
    A4, B4 = A2.copy(), B2.copy()
    A4[5, 7] = 1e-30
    fell_back, failed = run(A4, B4)
    assert fell_back == 1 and failed > 0
    # (5) the same entry with supports that overlap everywhere: nothing is zero, everything is large enough
    A5 = A.copy()
    A5[5, 7] = 1e-30
    assert run(A5, B) == (0, 0)


def test_predict_f64_int8_mode_against_oracle(ss, o):
    S, Yfull = _enzyme_like(o, seed=6)
    N, Nt = Yfull.shape
    names = [f"D{i:04d}" for i in range(N)]
    tn = [f"T{j:04d}" for j in range(Nt)]
    DT = ss.NamedArray(Yfull, (names, tn))
    for weighted, alpha in ((True, 0.2), (False, 0.35)):
        X = ss.featurize(ss.NamedArray(S, (names, names)), alpha, weighted)
        q = names[::9]
        A, B = ss.construct(DT, X, q)
        got = ss.predict((A, B), DT[q, tn], precision="f64_int8", clean=True)
        Ao, Bo, nn = o.construct_queries(Yfull, (names, tn), X.array, (names, X.names(2)), q)
        want = o.predict_dense(Ao, Bo, nn, q, tn)
        o.clean(want, Ao, nn, tn)
        assert relerr(got.array, want) < RTOL and np.array_equal(got.array == -99, want == -99)


def test_two_layer_nbi_recommender_shape(ss, o):
    """BASELINE config 5's form (scaled down): classical 2-layer NBI, no feature layer --
    F[s,t] = Y * U, U = (Y' ./ kt) * (Y ./ ks) (reference src/core.jl:446-466 on a graph without
    features), followed by per-user top-L."""
    from simspread_b200._lib import check
    rng = np.random.default_rng(55)
    users, items = 700, 420
    Y = (rng.random((users, items)) < 0.02).astype(float)
    Y[:, 7] = 0.0   # an item nobody has
    Y[11, :] = 0.0  # a user without items
    ctx = ss.Context.default()
    dY, R = ss.DMat.from_host(ctx, Y), ss.DMat(ctx, users, items)
    check(ss.lib().ss_predict_source(ctx.h, None, dY.h, R.h, 0))
    ks, kt = Y.sum(1), Y.sum(0)
    Wst = o._div_rows(Y, ks.astype(np.int64))
    U = o._div_rows(np.ascontiguousarray(Y.T), kt.astype(np.int64)) @ Wst
    want = Y @ U
    # literal reference path on the (users + items)^2 adjacency matrix
    n = users + items
    A = np.zeros((n, n))
    A[:users, users:] = Y
    A[users:, :users] = Y.T
    names = [f"n{i}" for i in range(n)]
    lit = o.predict_dense_single(A, names, names[:users], names[users:])
    assert relerr(want, lit) < 1e-12
    got = R.to_host()
    assert relerr(got, lit) < RTOL
    L = 20
    idx = ss.DIVec(ctx, L * users)
    check(ss.lib().ss_topl_rows(ctx.h, R.h, L, idx.h, None))
    order = np.stack([o.sortperm_rev(got[u])[:L] for u in range(users)])
    assert np.array_equal(idx.to_host().reshape(users, L), order)


# (40, 3000, 0.5): ~1500 targets per source -> several staging passes of the fused kernel (1024 items each) and the
# > 64-target tail loop of every co-rater; (3, 70000, 0.01): fewer sources than clusters, column tail of the row
@pytest.mark.parametrize("users,items,dens,L", [(300, 200, 0.05, 16), (1000, 2600, 0.01, 16), (64, 40000, 0.002, 32),
                                                (40, 3000, 0.5, 20), (3, 70001, 0.01, 8)])
def test_sparse_recommender_topl_against_dense_path(ss, o, users, items, dens, L):
    """Config-5 form at test scale: two-hop CSR expansion with L2-resident accumulators and fused
    top-L, against the dense chain (ss_predict_source) + ss_topl_rows and the oracle."""
    from simspread_b200._lib import check
    rng = np.random.default_rng(users + items)
    mask = rng.random((users, items)) < dens
    ctx = ss.Context.default()
    for weighted in (True, False):
        Y = np.where(mask, np.round(rng.random((users, items)) + 0.5, 3), 0.0) if weighted else mask.astype(float)
        Y[min(3, users - 1), :] = 0.0  # a user without items: all scores 0 -> first L columns
        idx, val = ss.recommend_topl(Y, L)
        ks = np.count_nonzero(Y, axis=1)
        kt = np.count_nonzero(Y, axis=0)
        U = o._div_rows(np.ascontiguousarray(Y.T), kt) @ o._div_rows(Y, ks)
        F = Y @ U
        want_val = -np.sort(-F, axis=1)[:, :L]
        got_val = np.take_along_axis(F, np.maximum(idx, 0), axis=1)
        assert idx.min() >= 0 and idx.max() < items
        # the selected scores are the L largest (element-wise within the FP64 tolerance) ...
        assert np.allclose(val, want_val, rtol=1e-12, atol=1e-300)
        assert np.allclose(got_val, want_val, rtol=1e-12, atol=1e-300)
        # ... in descending order, without duplicates
        assert np.all(val[:, :-1] >= val[:, 1:])
        assert all(len(set(r)) == L for r in idx)
        if weighted:  # no exact ties between non-zero scores: the order itself must match the reference order
            order = np.stack([o.sortperm_rev(F[u])[:L] for u in range(users)])
            # (the partial products are accumulated with RED.ADD.F64 in no fixed order, so scores that tie in
            # exact arithmetic may differ in the last bits: compare the order where the L + 1 best differ)
            top = -np.sort(-F, axis=1)[:, :L + 1]
            distinct = np.abs(np.diff(top, axis=1)) > 1e-9 * np.abs(top[:, :-1])
            rows_ok = distinct.all(axis=1) & (want_val[:, -1] > 0)
            assert (rows_ok.sum() > 0 or users < 10) and np.array_equal(idx[rows_ok], order[rows_ok])
        assert np.array_equal(idx[min(3, users - 1)], np.arange(L))  # all-zero row: stable order = first L columns


# ---------------------------------------------------------------------------------------------
# similarity computation fused with the threshold
# ---------------------------------------------------------------------------------------------


def test_jaccard_featurize_iris_golden(ss, o, iris):
    """The tutorial pipeline (docs/src/tutorial/fishers-flowers.jl:66, 97-99) from the raw descriptors:
    the fused kernel must reproduce featurize(S[test, train], 0.9, true) of the shipped iris.simmat."""
    names = iris["names"]
    D = ss.NamedArray(iris["F"], (names, ["sl", "sw", "pl", "pw"]))
    test, train = names[::10], [n for i, n in enumerate(names) if i % 10]
    S = o.jaccard_similarity(iris["F"], iris["F"])
    ri, ci = [names.index(t) for t in test], [names.index(t) for t in train]
    for weighted in (True, False):
        got = ss.jaccard_featurize(D, test, train, 0.9, weighted)
        want = o.cutoff(S[np.ix_(ri, ci)], 0.9, weighted)
        assert np.array_equal(got.array, want)
        assert got.names(2) == ["f" + t for t in train] and got.names(1) == test
        shipped = o.cutoff(iris["S"][np.ix_(ri, ci)], 0.9, weighted)  # the reference's own matrix, bit for bit
        assert np.array_equal(got.array, shipped)


@pytest.mark.parametrize("na,nb,d", [(1, 1, 1), (65, 130, 17), (300, 257, 64), (64, 64, 16)])
def test_jaccard_featurize_random_bit_exact(ss, o, na, nb, d):
    rng = np.random.default_rng(na * 1000 + nb + d)
    A, B = rng.random((na, d)), rng.random((nb, d))
    A[0, :] = 0.0
    B[0, :] = 0.0  # 0/0 -> distance 0 -> similarity 1
    ctx = ss.Context.default()
    from simspread_b200._lib import check
    dA, dB = ss.DMat.from_host(ctx, A), ss.DMat.from_host(ctx, B)
    for alpha, weighted in ((0.0, True), (0.55, True), (0.55, False)):
        X = ss.DMat(ctx, na, nb)
        check(ss.lib().ss_jaccard_featurize(ctx.h, dA.h, dB.h, alpha, int(weighted), X.h))
        want = o.cutoff(o.jaccard_similarity(A, B), alpha, weighted)
        got = X.to_host()
        assert np.array_equal(got, want)
        assert got[0, 0] == 1.0


@pytest.mark.parametrize("na,nb,words", [(3, 5, 1), (70, 129, 16), (256, 64, 32), (33, 31, 5)])
def test_tanimoto_bits_featurize_bit_exact(ss, o, na, nb, words):
    rng = np.random.default_rng(na + nb + words)
    FA = rng.integers(0, 2**63, size=(na, words), dtype=np.uint64) & rng.integers(0, 2**63, size=(na, words), dtype=np.uint64)
    FB = rng.integers(0, 2**63, size=(nb, words), dtype=np.uint64) & rng.integers(0, 2**63, size=(nb, words), dtype=np.uint64)
    FA[0] = 0
    FB[0] = 0
    T = o.tanimoto_bits(FA, FB)
    for alpha, weighted in ((0.0, True), (0.2, True), (0.2, False)):
        got = ss.tanimoto_featurize_bits(FA, FB, alpha, weighted)
        assert np.array_equal(got, o.cutoff(T, alpha, weighted))


# ---------------------------------------------------------------------------------------------
# AuROC / AuPRC of a list split into key ranges (the per-rank work of the multi-GPU form)
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("m,nseg", [(5000, 3), (100001, 8), (64, 5)])
def test_auc_key_range_segments_add_up(ss, o, m, nseg):
    """ss_auc_sort / _lower_bound / _segment_summary / _segment_integrate: the sorted list is cut at key splitters
    (equal keys never straddle a cut), every range is integrated on its own with what lies below it, and the signed
    partial areas add up to ss_auroc_auprc of the whole list and to the oracle."""
    import torch
    from simspread_b200._lib import check
    from simspread_b200.sharded import combine_segment_summaries, pick_splitters
    rng = np.random.default_rng(m + nseg)
    sc = np.round(rng.random(m), 3)  # many ties
    sc[:m // 20] = 0.0
    sc[m // 20:m // 16] = -0.0
    lb = (rng.random(m) < 0.1 + 0.6 * sc).astype(np.uint8)
    ctx = ss.Context.default()
    L = ss.lib()
    dev = torch.device("cuda", ctx.device)
    ts, tl = torch.from_numpy(sc).to(dev), torch.from_numpy(lb).to(dev)
    whole = (C.c_double * 2)()
    check(L.ss_auroc_auprc(ctx.h, C.c_void_p(tl.data_ptr()), C.c_void_p(ts.data_ptr()), m, whole))
    assert whole[0] == pytest.approx(o.AuROC(lb > 0, sc), rel=1e-12) and whole[1] == pytest.approx(o.AuPRC(lb > 0, sc), rel=1e-12)
    pk, pl = C.c_void_p(), C.c_void_p()
    check(L.ss_auc_sort(ctx.h, C.c_void_p(tl.data_ptr()), C.c_void_p(ts.data_ptr()), None, m, C.byref(pk), C.byref(pl)))
    keys = np.sort(o._isless_key(sc).astype(np.uint64))
    split = pick_splitters(keys[:: max(1, m // 40)], nseg)
    split[0] = keys[0]  # an empty first range
    cut = np.zeros(nseg - 1, dtype=np.int64)
    check(L.ss_auc_lower_bound(ctx.h, pk, m, split.ctypes.data, nseg - 1, cut.ctypes.data))
    assert np.array_equal(cut, np.searchsorted(keys, split, side="left"))
    bounds = np.concatenate([[0], np.maximum.accumulate(cut), [m]])
    # the segments are slices of the sorted arrays; copy them out (the sort buffers are reused below)
    from simspread_b200.sharded import _CudaView
    K = torch.as_tensor(_CudaView(pk.value, (m,), "<i8"), device=dev).clone()
    Lb = torch.as_tensor(_CudaView(pl.value, (m,), "|u1"), device=dev).clone()
    assert np.array_equal(K.cpu().numpy().view(np.uint64), keys)
    segs = [(K[bounds[i]:bounds[i + 1]].contiguous(), Lb[bounds[i]:bounds[i + 1]].contiguous()) for i in range(nseg)]
    sizes, summ = [], []
    for k_, l_ in segs:
        out = np.zeros(3, dtype=np.int64)
        check(L.ss_auc_segment_summary(ctx.h, C.c_void_p(k_.data_ptr()), C.c_void_p(l_.data_ptr()), k_.numel(), out.ctypes.data))
        sizes.append(k_.numel())
        summ.append(out.copy())
    assert sum(int(s_[0]) for s_ in summ) == int(lb.sum()) and sizes[0] == 0 and tuple(summ[0]) == (0, -1, 0)
    roc = pr = 0.0
    for r_, (k_, l_) in enumerate(segs):
        out = np.zeros(3, dtype=np.int64)
        check(L.ss_auc_segment_summary(ctx.h, C.c_void_p(k_.data_ptr()), C.c_void_p(l_.data_ptr()), k_.numel(), out.ctypes.data))
        g6 = np.array(combine_segment_summaries(sizes, summ, r_), dtype=np.int64)
        part = (C.c_double * 2)()
        check(L.ss_auc_segment_integrate(ctx.h, C.c_void_p(k_.data_ptr()), C.c_void_p(l_.data_ptr()), k_.numel(), g6.ctypes.data, part))
        roc += part[0]
        pr += part[1]
    assert abs(roc) == pytest.approx(whole[0], rel=1e-12) and abs(pr) == pytest.approx(whole[1], rel=1e-12)
