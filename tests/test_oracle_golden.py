"""Pins the CPU oracle (oracle/simspread_oracle.py) against every golden vector the reference's
own test-suite holds for the hot path.  CPU only."""
import math

import numpy as np
import pytest

from oracle import simspread_oracle as o


def test_k(kats):
    M = np.array(kats["k"]["M"], dtype=np.float64)
    assert o.k_node(0, M) == 0
    assert o.k_vec(M[0, :]) == 0
    assert o.k_mat(M).ravel().tolist() == kats["k"]["expect"]
    assert o.k_mat(M).shape == (4, 1)  # mapslices(dims=2) -> n x 1
    assert o.k_vec(np.array([0.0, -0.0, np.nan, 1.0])) == 2  # -0.0 is zero, NaN is not


def test_cutoff(kats):
    c = kats["cutoff"]
    x, y, z = c["x"], np.array(c["y"]).reshape(-1, 1), np.array(c["z"])
    for case in c["cases"]:
        a = case["alpha"]
        assert o.cutoff_scalar(x, a, False) == pytest.approx(case["x_bin"])
        assert o.cutoff_scalar(x, a, True) == pytest.approx(case["x_w"])
        yw = y.ravel() if case["y_w"] == "y" else np.array(case["y_w"], dtype=float)
        zw = z if case["z_w"] == "z" else np.array(case["z_w"], dtype=float)
        np.testing.assert_allclose(o.cutoff(y, a, False).ravel(), case["y_bin"])
        np.testing.assert_allclose(o.cutoff(y, a, True).ravel(), yw)
        np.testing.assert_allclose(o.cutoff(z, a, False), case["z_bin"])
        np.testing.assert_allclose(o.cutoff(z, a, True), zw)
    # inclusive threshold, NaN -> 0, argument not mutated by cutoff!
    assert o.cutoff_scalar(0.5, 0.5) == 1.0
    assert o.cutoff(np.array([np.nan]), 0.0)[0] == 0.0
    zz = z.copy()
    o.cutoff_inplace(zz, 0.5, False)
    assert np.array_equal(zz, z)


def test_featurize(kats):
    f = kats["featurize"]
    M0 = np.array(f["M0"])
    b, r, c = o.featurize(M0, f["names"], f["names"], f["alpha"], False)
    w, _, _ = o.featurize(M0, f["names"], f["names"], f["alpha"], True)
    assert np.array_equal(b, np.array(f["bin"], dtype=float))
    assert np.array_equal(w, np.array(f["w"], dtype=float))
    assert c == f["colnames"] and r == f["names"]


def test_construct_order_and_errors(kats):
    c = kats["construct"]
    X, y = np.array(c["X"], dtype=float), np.array(c["y"], dtype=float)
    A, B, names = o.construct_queries(y, (c["xrows"], c["ycols"]), X, (c["xrows"], c["xcols"]),
                                      c["queries"])
    assert names == c["names"]
    assert A.shape == (7, 7) and np.array_equal(A, A.T)
    assert not B[0, :].any() and not B[:, 0].any()
    assert np.array_equal(B[1:, 1:], A[1:, 1:])
    with pytest.raises(AssertionError, match=c["err_same_names"]):
        o.construct_queries(y, (c["xrows"], c["ycols"]), X, (c["xrows"], c["xrows"]), c["queries"])
    with pytest.raises(AssertionError, match=c["err_rows"]):
        o.construct_queries(y, (c["xrows"], c["ycols"]), X[:2], (c["xrows"][:2], c["xcols"]),
                            c["queries"])


def test_spread(kats):
    s = kats["spread"]
    got, want = o.spread(np.array(s["M"], dtype=float)), np.array(s["W"])
    # Julia `isapprox` on arrays is norm-based: |x-y| <= rtol*max(|x|,|y|)
    assert np.linalg.norm(got - want) <= s["rtol"] * max(np.linalg.norm(got), np.linalg.norm(want))
    assert np.array_equal(got, np.array([[1, 0, 0], [.5, .5, 0], [1 / 3, 1 / 3, 1 / 3]]))
    W = o.spread(np.zeros((3, 3)))
    assert np.array_equal(W, np.zeros((3, 3)))  # 0/0 = NaN -> 0


def test_predict_kat_exact(kats):
    p = kats["predict"]
    A, B = np.array(p["A"], dtype=float), np.array(p["B"], dtype=float)
    yhat = o.predict_dense(A, B, p["names"], p["rows"], p["cols"])
    assert np.array_equal(yhat, np.array(p["yhat"]))  # exact ==, as in the reference test
    # block-reduced form gives the same known answer
    Xq, Xs, Y = A[0:1, 4:7], A[1:4, 4:7], A[1:4, 7:9]
    assert np.array_equal(o.predict_blocks_query(Xq, Xs, Y), np.array(p["yhat"]))


def test_clean(kats):
    c = kats["clean"]
    yhat = np.array(c["yhat"], dtype=float)
    o.clean(yhat, np.array(c["A"], dtype=float), c["names"], c["targets"])
    assert np.array_equal(yhat, np.array(c["expect"], dtype=float))


def test_save(kats):
    s = kats["save"]
    y = np.array(s["y"])
    assert o.save_rows(y, y, s["rows"], s["cols"]) == s["files"]["save1"]
    assert o.save_rows(y, y, s["rows"], s["cols"], delimiter=" ") == s["files"]["save2"]
    assert o.save_rows(y, y, s["rows"], s["cols"], fold=1) == s["files"]["save3"]
    assert o.save_rows(y, y, s["rows"], s["cols"], fold=1, delimiter=" ") == s["files"]["save4"]


def test_atL(kats):
    a = kats["atL"]
    for L, v in a["recall"].items():
        assert o.recallatL_grouped(a["y"], a["yhat"], a["grouping"], int(L)) == pytest.approx(v)
    for L, v in a["precision"].items():
        assert o.precisionatL_grouped(a["y"], a["yhat"], a["grouping"], int(L)) == pytest.approx(v)
    with pytest.raises(AssertionError):
        o.recallatL(a["y"], a["yhat"], 10)  # strict length > L
    # all-negative group -> NaN propagates through the grouped mean
    assert math.isnan(o.recallatL_grouped([0, 0, 1, 0], [1, 2, 3, 4], [1, 1, 2, 2], 1))


def test_confusion_scalars(kats):
    c = kats["confusion"]
    t = tuple(c["tn_fp_fn_tp"])
    assert o.roc_int(c["y"], c["yhat"]) == t
    assert o.f1score(*t) == pytest.approx(c["f1"])
    assert o.mcc(*t) == pytest.approx(c["mcc"])
    assert o.accuracy(*t) == pytest.approx(c["acc"])
    assert o.balancedaccuracy(*t) == pytest.approx(c["bacc"])
    assert o.recall(*t) == pytest.approx(c["recall"])
    assert o.precision(*t) == pytest.approx(c["precision"])


def test_mcc_limits(kats):
    m = kats["mcc_limits"]
    y, yhat = m["y"], m["yhat"]
    ref = o.mcc_eps(5, 5)
    assert o.mcc(*o.roc_int(y, np.ones(10, int))) - ref < m["tol"]
    assert o.mcc(*o.roc_int(y, np.zeros(10, int))) - ref < m["tol"]
    assert o.mcc(*o.roc_int(np.ones(10, int), yhat)) - ref < m["tol"]
    assert o.mcc(*o.roc_int(np.zeros(10, int), yhat)) - ref < m["tol"]


# ---- restatement self-consistency (no reference golden exists: "parity unpinned") -------------


def test_auroc_auprc_quirks():
    # SURVEY.md App. A item 16 (hand-derived from src/performance.jl:53-61,78-86 + MLBase/Trapz)
    assert o.AuROC([1, 0, 1, 0, 0], [.9, .9, .7, .1, .1]) == pytest.approx(2 / 3)
    assert o.AuROC([1, 0, 1], [.5, .5, .5]) == 0.0
    assert o.AuPRC([1, 0, 1, 0], [.9, .8, .7, .1]) == pytest.approx(0.2916666666666667)


def test_auroc_matches_bruteforce():
    rng = np.random.default_rng(1)
    y = rng.random(300) < 0.3
    yhat = np.round(rng.random(300), 2)  # many ties
    thr = np.unique(yhat)
    tp = np.array([np.sum(y & (yhat >= t)) for t in thr])
    fp = np.array([np.sum(~y & (yhat >= t)) for t in thr])
    P, N = y.sum(), (~y).sum()
    x, v = fp / N, tp / P
    brute = abs(np.sum((x[1:] - x[:-1]) * (v[1:] + v[:-1]) / 2))
    assert o.AuROC(y, yhat) == pytest.approx(brute, rel=1e-13)
    r, p = tp / P, tp / (tp + fp)
    brute = abs(np.sum((r[1:] - r[:-1]) * (p[1:] + p[:-1]) / 2))
    assert o.AuPRC(y, yhat) == pytest.approx(brute, rel=1e-13)


def test_sortperm_rev_stable():
    v = np.array([1.0, 3.0, 3.0, -0.0, 0.0, np.nan, 2.0])
    assert o.sortperm_rev(v).tolist() == [5, 1, 2, 6, 0, 4, 3]


@pytest.mark.parametrize("weighted", [False, True])
def test_block_form_equals_dense_form(weighted):
    """the block-reduced chain equals the literal n x n path."""
    rng = np.random.default_rng(7)
    nq, ns, nf, nt = 7, 13, 13, 5
    S = np.round(rng.random((nq + ns, nf)), 3)
    X = o.cutoff(S, 0.4, weighted)
    X[nq + 2, :] = 0.0  # an isolated-feature source
    Y = (rng.random((ns, nt)) < 0.3).astype(float)
    Y[:, 2] = 0.0  # a target without edges
    Y[2, :] = 0.0  # with the line above: an isolated source (k = 0 -> 0/0 -> 0)
    Xq, Xs = X[:nq], X[nq:]
    A = o._assemble4(Xq, Xs, Y)
    B = A.copy()
    B[:nq, :] = 0
    B[:, :nq] = 0
    names = [f"n{i}" for i in range(A.shape[0])]
    Fq = o.predict_dense(A, B, names, names[:nq], names[nq + ns + nf:])
    np.testing.assert_allclose(o.predict_blocks_query(Xq, Xs, Y), Fq, rtol=1e-13, atol=1e-300)
    # 3-layer graph / source rows (src/core.jl:446-466)
    A3 = A[nq:, nq:]
    n3 = names[nq:]
    Fs = o.predict_dense_single(A3, n3, n3[:ns], n3[ns + nf:])
    np.testing.assert_allclose(o.predict_blocks_source(Xs, Y), Fs, rtol=1e-13, atol=1e-300)
    # clean! in block form
    R1 = Fq.copy()
    o.clean(R1, A, names, names[nq + ns + nf:])
    R2 = Fq.copy()
    o.clean_blocks(R2, o.degrees_blocks(Xs, Y)[2])
    assert np.array_equal(R1, R2) and (R1[:, 2] == -99).all()


def test_iris_fixture_runs(iris):
    S, C, names = iris["S"], iris["C"], iris["names"]
    X, xr, xc = o.featurize(S, names, names, 0.9, True)
    q = names[::10]
    A, B, nn = o.construct_queries(C, (names, iris["classes"]), X, (xr, xc), q)
    assert A.shape[0] == 15 + 135 + 135 + 3
    yhat = o.predict_dense(A, B, nn, q, iris["classes"])
    qi = [names.index(x) for x in q]
    si = [i for i in range(150) if i not in qi]
    blk = o.predict_blocks_query(X[np.ix_(qi, si)], X[np.ix_(si, si)], C[si])
    np.testing.assert_allclose(blk, yhat, rtol=1e-13, atol=1e-300)
    assert 0.9 < o.AuROC(C[qi].ravel() > 0, yhat.ravel()) <= 1.0


def test_jaccard_similarity_pinned_by_iris_simmat(iris):
    """the tutorial's `1 .- pairwise(Jaccard(), X, dims=1)` (docs/src/tutorial/fishers-flowers.jl:66).
    docs/src/tutorial/data/iris.simmat is that matrix for iris.features: the restatement reproduces all
    22 500 shipped values bit for bit."""
    S = o.jaccard_similarity(iris["F"], iris["F"])
    assert S.shape == (150, 150)
    assert np.array_equal(S, iris["S"])
    assert np.array_equal(np.diag(S), np.ones(150))
    assert o.jaccard_similarity(np.zeros((1, 3)), np.zeros((2, 3))).tolist() == [[1.0, 1.0]]  # 0/0: distance 0


def test_tanimoto_bits_matches_jaccard_of_indicator_vectors():
    rng = np.random.default_rng(3)
    FA = rng.integers(0, 2**63, size=(7, 3), dtype=np.uint64) & rng.integers(0, 2**63, size=(7, 3), dtype=np.uint64)
    FB = rng.integers(0, 2**63, size=(5, 3), dtype=np.uint64)
    FA[2] = 0
    FB[1] = 0
    T = o.tanimoto_bits(FA, FB)
    ba = np.unpackbits(FA.view(np.uint8), axis=1).astype(float)
    bb = np.unpackbits(FB.view(np.uint8), axis=1).astype(float)
    J = o.jaccard_similarity(ba, bb)
    assert np.allclose(T, J, rtol=0, atol=2.3e-16) and T[2, 1] == 1.0  # J goes through 1 - (1 - q)


def test_auroc_auprc_restatement_against_scikit_learn_counts():
    """AuROC / AuPRC have no golden in the reference (placeholders at test/runtests.jl:210-217), so the restatement of
    MLBase.roc + Trapz.trapz is cross-checked against an independent implementation of the confusion counts:
    scikit-learn's roc_curve / precision_recall_curve give (fpr, tpr) / (precision, recall) at every unique threshold;
    the reference's curve is exactly those points WITHOUT sklearn's synthetic anchors ((0,0) for ROC, (recall 0,
    precision 1) for PR) -- src/performance.jl:53-61, 78-86, SURVEY App. A.16."""
    sk = pytest.importorskip("sklearn.metrics")
    trapezoid = getattr(np, "trapezoid", None) or np.trapz
    rng = np.random.default_rng(0)
    for trial in range(6):
        n = 1500 + 100 * trial
        s = np.round(rng.random(n), 2 if trial % 2 else 6)  # heavy ties / almost none
        y = rng.random(n) < 0.1 + 0.5 * s
        fpr, tpr, _ = sk.roc_curve(y, s, drop_intermediate=False)
        assert o.AuROC(y, s) == pytest.approx(abs(trapezoid(tpr[1:], fpr[1:])), rel=1e-12)
        prec, rec, _ = sk.precision_recall_curve(y, s)
        assert o.AuPRC(y, s) == pytest.approx(abs(trapezoid(prec[:-1], rec[:-1])), rel=1e-12)
    # the documented consequence of the missing anchors: 2/3 instead of the textbook 0.75
    assert o.AuROC(np.array([1, 0, 1, 0, 0], bool), np.array([.9, .9, .7, .1, .1])) == pytest.approx(2 / 3, rel=1e-12)


@pytest.mark.parametrize("weighted", [False, True])
def test_two_layer_ordered_oracle_pins_scipy(weighted):
    """The oracle of the sparse recommender form (BASELINE config 5) is the reference's association A * (W * W)
    (src/core.jl:456) with every sum in ascending index order.  Its fast form uses scipy.sparse CSR x CSR products; this
    pins that they add in exactly that order (bit-equal to literal Python loops) and that the result agrees with the
    dense block form and the literal n x n path within the FP64 tolerance -- but NOT bit for bit (BLAS order), which is
    why the bit-exact top-L tests need the ordered form."""
    rng = np.random.default_rng(4 + int(weighted))
    mask = rng.random((60, 45)) < 0.15
    Y = np.where(mask, np.round(rng.random((60, 45)) + 0.5, 3), 0.0) if weighted else mask.astype(float)
    Y[7, :] = 0.0
    Y[:, 3] = 0.0
    F_loops = o.two_layer_scores_loops(Y)
    F_sp, U_sp = o.two_layer_scores_sparse(Y)
    assert np.array_equal(F_sp.toarray(), F_loops)
    assert np.array_equal(U_sp.toarray(), o.two_layer_transfer_loops(Y))
    F_rows, _ = o.two_layer_scores_sparse(Y, rows=[5, 7, 41])
    assert np.array_equal(F_rows.toarray(), F_loops[[5, 7, 41]])
    n = 60 + 45
    A = np.zeros((n, n))
    A[:60, 60:] = Y
    A[60:, :60] = Y.T
    names = [f"n{i}" for i in range(n)]
    lit = o.predict_dense_single(A, names, names[:60], names[60:])
    nz = lit != 0
    assert np.max(np.abs(F_loops[nz] - lit[nz]) / lit[nz]) < 1e-13 and not F_loops[~nz].any()
    idx, val = o.recommend_topl(Y, 5)
    assert np.array_equal(idx[0], o.sortperm_rev(F_loops[0])[:5]) and np.array_equal(val[0], F_loops[0][idx[0]])
    assert np.array_equal(idx[7], np.arange(5))   # a source without items: all scores 0 -> the first columns
