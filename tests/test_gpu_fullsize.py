"""`-m gpu` tests at BASELINE.json's FULL sizes, where the oracle cannot run: the CUDA path is checked through
size-independent properties of the domain -- checksums of checksums (row / column sums of a product are products of
sums: the linearity of predict in its operands), resource conservation of the spreading step, idempotence of the
threshold, degree sum rules, determinism of row slabs, sortedness and permutation invariance of the ranking
metrics -- plus sampled entries recomputed independently with torch (FP64) at the north-star tolerance.

Inputs are generated on the device (they do not fit the host comfortably); torch is the checker here, never the
thing measured.  Tolerance for FP64 scores: 1e-12 relative (north_star)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-12


@pytest.fixture(scope="module")
def env():
    import torch
    import simspread_b200 as ss
    from simspread_b200._lib import check
    ss.build()
    ctx = ss.Context.default()
    dev = torch.device("cuda", ctx.device)
    free, total = torch.cuda.mem_get_info(dev)
    return {"ss": ss, "check": check, "ctx": ctx, "torch": torch, "dev": dev, "L": ss.lib(), "free_gb": free / 2**30}


def colmajor(torch, rows, cols, dev):
    ld = (rows + 15) // 16 * 16
    return torch.zeros((cols, ld), dtype=torch.float64, device=dev), ld


def fill(torch, buf, rows, seed, kind, p=0.05):
    g = torch.Generator(device=buf.device)
    g.manual_seed(seed)
    step = max(1, (1 << 27) // buf.shape[1])
    for c0 in range(0, buf.shape[0], step):
        blk = buf[c0:c0 + step, :rows]
        u = torch.rand(blk.shape, generator=g, device=buf.device, dtype=torch.float64)
        blk.copy_(torch.round(u * 1e6) / 1e6 if kind == "uniform6" else (u < p).to(torch.float64))


def relmax(torch, got, want):
    return float(((got - want).abs() / want.abs().clamp_min(1e-300)).max().item())


def test_c4_full_size_chain_properties(env):
    """BASELINE config 4 (100k queries x 20k sources/features x 50k targets, FP64): spread + predict + clean!."""
    ss, check, ctx, torch, dev, L = (env[k] for k in ("ss", "check", "ctx", "torch", "dev", "L"))
    if env["free_gb"] < 120:
        pytest.skip("needs ~100 GB of device memory")
    nq, ns, nf, nt = 100_000, 20_000, 20_000, 50_000
    bXq, ldq = colmajor(torch, nq, nf, dev)
    bXs, lds = colmajor(torch, ns, nf, dev)
    bY, ldy = colmajor(torch, ns, nt, dev)
    bR, ldr = colmajor(torch, nq, nt, dev)
    fill(torch, bXq, nq, 1, "uniform6")
    fill(torch, bXs, ns, 2, "uniform6")
    fill(torch, bY, ns, 3, "bernoulli", 0.05)
    bY[7, :] = 0.0      # a target nobody has: clean! must flag its column
    bXs[:, 11] = 0.0    # a source without features (it still has targets)
    bXs[13, :] = 0.0    # a feature nobody has: kf = 0 -> its row of T is 0, not NaN
    torch.cuda.synchronize()  # the library has its own stream: torch's writes must have landed before it reads them
    mXq, mXs = ss.DMat.wrap(ctx, bXq.data_ptr(), nq, nf, ldq), ss.DMat.wrap(ctx, bXs.data_ptr(), ns, nf, lds)
    mY, mR = ss.DMat.wrap(ctx, bY.data_ptr(), ns, nt, ldy), ss.DMat.wrap(ctx, bR.data_ptr(), nq, nt, ldr)
    kt = ss.DIVec(ctx, nt)
    check(L.ss_predict_query(ctx.h, mXq.h, mXs.h, mY.h, mR.h, ss._lib.SS_PREDICT_CLEAN, kt.h))
    torch.cuda.synchronize()
    Xq, Xs, Y, R = bXq[:, :nq], bXs[:, :ns], bY[:, :ns], bR[:, :nq]  # torch views, TRANSPOSED (row = column of the matrix)
    # degrees: independent recount, and the sum rules sum(ks) = sum(kf) + sum(kt) = nnz(Xs) + nnz(Y)
    ks_ = (Xs != 0).sum(0) + (Y != 0).sum(0)
    kf_ = (Xs != 0).sum(1)
    kt_ = (Y != 0).sum(1)
    assert np.array_equal(kt.to_host(), kt_.cpu().numpy().astype(np.int32))
    assert int(ks_.sum()) == int(kf_.sum()) + int(kt_.sum())
    assert int(kt_[7]) == 0 and int(kf_[13]) == 0
    # clean!: exactly the columns of degree-0 targets are -99, everything else is a finite non-negative score
    flagged = (R == -99.0).all(1)
    assert bool(torch.equal(flagged, kt_ == 0)) and int(flagged.sum()) >= 1
    assert bool(torch.isfinite(R).all()) and float(R.amin(1)[~flagged].min()) >= 0.0
    livemask = (~flagged).to(torch.float64)
    # checksum of checksums (linearity): with w[s] = 1/ks[s], a[f] = 1/kf[f] (0 where the degree is 0)
    #   row sums    R 1 = Xq (T 1),          T 1 = a .* (Xs' (w .* (Y 1)))           (over the live targets)
    #   column sums 1' R = (1' Xq) T  ->  checked for a sample of columns through T's definition
    w = torch.where(ks_ > 0, 1.0 / ks_.to(torch.float64), torch.zeros(ns, dtype=torch.float64, device=dev))
    a = torch.where(kf_ > 0, 1.0 / kf_.to(torch.float64), torch.zeros(nf, dtype=torch.float64, device=dev))
    y1 = torch.mv(Y.T, livemask)                   # Y 1 over live targets (length ns)
    t1 = a * torch.mv(Xs, w * y1)                  # T 1 (length nf); Xs view is (nf, ns) = Xs'
    want_rows = torch.mv(Xq.T, t1)                 # Xq (T 1)
    got_rows = torch.mv(R.T, livemask)             # R 1 over the live targets, no 40 GB temporary
    assert relmax(torch, got_rows, want_rows) < 5e-12  # sums of 5e4 non-negative terms, two summation orders
    xq1 = Xq.sum(1)                                # 1' Xq (length nf)
    g = torch.Generator(device="cpu")
    g.manual_seed(5)
    cols = torch.randint(0, nt, (48,), generator=g).to(dev)
    cols = cols[~flagged[cols]]
    Tc = a[:, None] * (Xs @ (w[:, None] * Y[cols].T))   # T[:, cols]  (nf x 48), torch / cuBLAS FP64
    assert relmax(torch, R[cols].sum(1), xq1 @ Tc) < 5e-12
    # sampled entries recomputed independently: R[q, t] = Xq[q, :] . T[:, t]
    rows = torch.randint(0, nq, (cols.numel(),), generator=g).to(dev)
    want = (Xq[:, rows] * Tc).sum(0)
    assert relmax(torch, R[cols, rows], want) < RTOL
    # resource conservation of one spreading step: every source with neighbours hands on exactly its resource,
    # sum_t Wst[s, t] = ky[s] / ks[s]  -> sum over the T built from it: sum_f kf[f] T[f, t] = sum_s (Xs 1)[s] w[s] Y[s, t]
    lhs = (kf_.to(torch.float64)[:, None] * Tc).sum(0)
    rhs = ((Xs.sum(0) * w)[:, None] * Y[cols].T).sum(0)
    assert relmax(torch, lhs, rhs) < 5e-12
    # SURVEY 8(d) protocol for this size: >= 1e4 sampled entries recomputed on the HOST in extended precision (x87
    # long double, 64-bit mantissa) from the reference's own formula -- W[f,s] = fl(Xs[s,f] / kf[f]),
    # W[s,t] = fl(Y[s,t] / ks[s]) rounded to Float64 as `G ./ k(G)` does (src/core.jl:366), the sums of
    # (W*W)[f,t] and of A * (W*W) exact to ~1e-19 -- against the FP64 result of the GPU at the north-star 1e-12.
    ncol_s, nrow_s = 8, 1280
    live = cols.cpu().numpy()
    live = live[np.sort(np.unique(live, return_index=True)[1])][:ncol_s]
    assert len(live) == ncol_s
    rs = np.sort(np.random.default_rng(17).choice(nq, size=nrow_s, replace=False))
    Xs_h = Xs.cpu().numpy()                                   # (nf, ns): row f = feature f over the sources
    ks_h, kf_h = ks_.cpu().numpy(), kf_.cpu().numpy()
    Y_h = Y[torch.from_numpy(live).to(dev)].cpu().numpy()     # (8, ns)
    Xq_h = Xq[:, torch.from_numpy(rs).to(dev)].cpu().numpy()  # (nf, 1280)
    got_h = R[torch.from_numpy(live).to(dev)][:, torch.from_numpy(rs).to(dev)].cpu().numpy()  # (8, 1280)
    worst = 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        for j in range(ncol_s):
            src = np.flatnonzero(Y_h[j])
            w_st = Y_h[j, src] / ks_h[src].astype(np.float64)                 # W[s,t], Float64 as in the reference
            w_fs = np.where(kf_h[:, None] > 0, Xs_h[:, src] / kf_h[:, None].astype(np.float64), 0.0)  # W[f,s]
            t_col = (w_fs.astype(np.longdouble) * w_st.astype(np.longdouble)[None, :]).sum(axis=1)     # (W*W)[f,t]
            want_ld = (Xq_h.astype(np.longdouble) * t_col[:, None]).sum(axis=0)                          # F[q,t]
            rel = np.abs(got_h[j].astype(np.longdouble) - want_ld) / np.maximum(np.abs(want_ld), np.longdouble(1e-300))
            worst = max(worst, float(rel.max()))
    assert ncol_s * nrow_s >= 10_000 and worst < RTOL, worst
    del Xs_h, Xq_h
    # determinism / slab independence: the first 1000 query rows computed on their own are bit-identical
    bR2, ldr2 = colmajor(torch, 1000, nt, dev)
    mXq2 = ss.DMat.wrap(ctx, bXq.data_ptr(), 1000, nf, ldq)
    mR2 = ss.DMat.wrap(ctx, bR2.data_ptr(), 1000, nt, ldr2)
    check(L.ss_predict_query(ctx.h, mXq2.h, mXs.h, mY.h, mR2.h, ss._lib.SS_PREDICT_CLEAN, None))
    torch.cuda.synchronize()
    assert bool(torch.equal(bR2[:, :1000], R[:, :1000]))


def test_c4_full_size_featurize_degrees_csr(env):
    """The threshold / degree / CSR kernels on the C4-sized similarity block (120k x 20k, 19 GB)."""
    ss, check, ctx, torch, dev, L = (env[k] for k in ("ss", "check", "ctx", "torch", "dev", "L"))
    if env["free_gb"] < 80:
        pytest.skip("needs ~60 GB of device memory")
    n, m, alpha = 120_000, 20_000, 0.93
    bS, ld = colmajor(torch, n, m, dev)
    fill(torch, bS, n, 9, "uniform6")
    bS[3, 5] = float("nan")   # NaN >= alpha is false -> 0
    bS[4, 6] = alpha          # inclusive threshold
    bX, _ = colmajor(torch, n, m, dev)
    torch.cuda.synchronize()  # torch's fills (other stream) before the library reads them
    mS, mX = ss.DMat.wrap(ctx, bS.data_ptr(), n, m, ld), ss.DMat.wrap(ctx, bX.data_ptr(), n, m, ld)
    check(L.ss_featurize(ctx.h, mS.h, alpha, 1, mX.h))
    torch.cuda.synchronize()
    S, X = bS[:, :n], bX[:, :n]
    keep = S >= alpha
    assert bool(torch.equal(X != 0, keep)) and float(X[3, 5]) == 0.0 and float(X[4, 6]) == alpha
    assert bool(torch.equal(X[keep], S[keep]))               # weighted: kept entries unchanged, bit for bit
    check(L.ss_featurize(ctx.h, mX.h, alpha, 1, mX.h))          # idempotent (featurize! in place)
    torch.cuda.synchronize()
    assert bool(torch.equal(X[keep], S[keep])) and int((X != 0).sum()) == int(keep.sum())
    # degrees of the featurized block: row / column counts against torch, and sum(rows) == sum(cols) == nnz
    kr, kc = ss.DIVec(ctx, n), ss.DIVec(ctx, m)
    check(L.ss_degrees(ctx.h, None, mX.h, kr.h, None, kc.h))   # row and column non-zero counts of the block
    krh, kch = kr.to_host().astype(np.int64), kc.to_host().astype(np.int64)
    assert np.array_equal(krh, keep.sum(0).cpu().numpy()) and np.array_equal(kch, keep.sum(1).cpu().numpy())
    assert krh.sum() == kch.sum() == int(keep.sum())
    # CSR straight from the raw similarities: same edge count, row extents = row degrees, ascending columns
    h = C.c_void_p()
    check(L.ss_featurize_csr(ctx.h, mS.h, alpha, 1, C.byref(h)))
    csr = ss.DCsr(ctx, h)
    assert csr.nnz == int(keep.sum()) and (csr.rows, csr.cols) == (n, m)
    rp, ci, va = csr.to_host()
    assert np.array_equal(np.diff(rp.astype(np.int64)), krh)
    inner = np.ones(csr.nnz, dtype=bool)
    inner[rp[1:-1][rp[1:-1] < csr.nnz]] = False            # first entry of every row
    assert np.all(np.diff(ci.astype(np.int64))[inner[1:]] > 0)
    r0 = 77_777
    assert np.array_equal(ci[rp[r0]:rp[r0 + 1]], np.flatnonzero(keep[:, r0].cpu().numpy()))
    assert np.array_equal(va[rp[r0]:rp[r0 + 1]], S[:, r0][keep[:, r0]].cpu().numpy())


def test_full_size_ranking_metric_properties(env):
    """AuROC / AuPRC / top-L on 10^9 scores: invariance under a permutation of the pairs and under a strictly
    increasing transform of the scores, the perfect / inverted rankings, sortedness of the top-L lists."""
    ss, check, ctx, torch, dev, L = (env[k] for k in ("ss", "check", "ctx", "torch", "dev", "L"))
    if env["free_gb"] < 80:
        pytest.skip("needs ~50 GB of device memory")
    M = 1_000_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(21)
    sc = torch.round(torch.rand(M, dtype=torch.float64, device=dev, generator=g) * 1e7) / 1e7  # ~10^7 distinct values: ties
    lb = (torch.rand(M, device=dev, generator=g) < 0.02 + 0.05 * sc).to(torch.uint8)

    def auc(labels, scores):
        out = (C.c_double * 2)()
        torch.cuda.synchronize()
        check(L.ss_auroc_auprc(ctx.h, C.c_void_p(labels.data_ptr()), C.c_void_p(scores.data_ptr()), scores.numel(), out))
        return out[0], out[1]

    base = auc(lb, sc)
    assert 0.5 < base[0] < 1.0 and 0.0 < base[1] < 1.0
    # a strictly increasing transform keeps every comparison: identical curve, bit for bit
    assert auc(lb, sc * 3.0 + 1.0) == base
    # reversing the order of the pairs changes nothing (ties are grouped by value, not by position)
    assert auc(lb.flip(0).contiguous(), sc.flip(0).contiguous()) == pytest.approx(base, rel=1e-12)
    del sc
    # perfect and inverted rankings of the same labels
    perfect = lb.to(torch.float64)
    a_perf = auc(lb, perfect)
    assert a_perf[0] == pytest.approx(1.0, rel=1e-12)
    a_inv = auc(lb, 1.0 - perfect)
    assert a_inv[0] == pytest.approx(0.0, abs=1e-12)
    del perfect, lb
    # top-L over a 100k x 5k score block: lists are sorted, hold the row maximum first, and no entry outside beats the last
    rows, cols, Ltop = 100_000, 5_000, 20
    bR, ldr = colmajor(torch, rows, cols, dev)
    fill(torch, bR, rows, 31, "uniform6")
    mR = ss.DMat.wrap(ctx, bR.data_ptr(), rows, cols, ldr)
    torch.cuda.synchronize()
    idx = ss.DIVec(ctx, Ltop * rows)
    check(L.ss_topl_rows(ctx.h, mR.h, Ltop, idx.h, None))
    ih = torch.from_numpy(idx.to_host().reshape(rows, Ltop).astype(np.int64)).to(dev)
    R = bR[:, :rows].T                                  # (rows, cols) view
    vals = torch.gather(R, 1, ih)
    assert bool((vals[:, :-1] >= vals[:, 1:]).all())
    assert bool(torch.equal(vals[:, 0], R.max(1).values))
    ties_in_order = (vals[:, :-1] > vals[:, 1:]) | (ih[:, :-1] < ih[:, 1:])   # equal scores: ascending column
    assert bool(ties_in_order.all())
    kth = torch.topk(R[:2000], Ltop, dim=1).values
    assert bool(torch.equal(kth, vals[:2000]))


@pytest.mark.parametrize("degrees,weighted", [("poisson", False), ("poisson", True), ("pareto", False)])
def test_c5_full_size_graph_user_slice_against_scipy(env, degrees, weighted):
    """BASELINE config 5 at FULL graph size (2M users x 500k items, ~1e8 edges): the transfer matrix U of the whole
    graph is built on the device and a slice of the users is ranked; sampled users are recomputed with scipy.sparse
    (CSR x CSR products add in the same ascending order, oracle/simspread_oracle.py): the top-20 columns must equal
    `sortperm(rev=true)[1:20]` row for row and the scores bit for bit -- binary graphs tie everywhere, so this is the
    north star's "bit-exact top-k order" where it is hardest -- and a second run must return identical bits.
    The Pareto-degree variant (users with up to 20 000 items) has a transfer matrix that approaches items x items: it is
    declined by the materialised form, `ss_recommend_topl` takes the two-hop expansion kernel (unordered FP64 sums), and
    the check is the FP64 tolerance plus the order wherever the scores differ."""
    import os
    import sys
    import scipy.sparse as sp
    from oracle import simspread_oracle as o
    ss, check, ctx, torch, dev, L = (env[k] for k in ("ss", "check", "ctx", "torch", "dev", "L"))
    if env["free_gb"] < 100:
        pytest.skip("needs ~70 GB of device memory")
    import gc
    gc.collect()
    torch.cuda.empty_cache()  # the C4 tests of this module leave tens of GB in torch's caching allocator
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from c5_graph import make_graph, wrap
    ns, nt, topl = 2_000_000, 500_000, 20
    G = make_graph(ns, nt, 1e-4, dev, degrees, weighted)
    hY, hYT = wrap(L, ctx, check, G)
    exact = degrees == "poisson"
    users = 20_000 if exact else 2_000  # a slice of the users; U is that of the whole graph
    idx = torch.full((ns, topl), -2, dtype=torch.int32, device=dev)
    val = torch.zeros((ns, topl), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()  # torch's fills must not land after the library (own stream) has written its results
    vi, vm = C.c_void_p(), C.c_void_p()
    check(L.ss_ivec_wrap(ctx.h, C.c_void_p(idx.data_ptr()), ns * topl, C.byref(vi)))
    check(L.ss_mat_wrap(ctx.h, C.c_void_p(val.data_ptr()), topl, ns, topl, C.byref(vm)))
    s0 = 1_000_000
    if exact:
        hU = C.c_void_p()
        check(L.ss_transfer_build(ctx.h, hY, hYT, C.byref(hU)))
        try:
            check(L.ss_recommend_topl_transfer(ctx.h, hY, hU, topl, s0, s0 + users, vi, vm))
            i1, v1 = idx[s0:s0 + users].clone(), val[s0:s0 + users].clone()
            idx[s0:s0 + users] = -2
            check(L.ss_recommend_topl_transfer(ctx.h, hY, hU, topl, s0, s0 + users, vi, vm))
            assert torch.equal(i1, idx[s0:s0 + users]) and torch.equal(v1.view(torch.int64), val[s0:s0 + users].view(torch.int64))
        finally:
            check(L.ss_transfer_destroy(hU))
    else:
        check(L.ss_recommend_topl(ctx.h, hY, hYT, topl, s0, s0 + users, vi, vm))
        i1, v1 = idx[s0:s0 + users].clone(), val[s0:s0 + users].clone()
    assert int(idx[:s0].max()) == -2 and int(idx[s0 + users:].max()) == -2   # rows outside the range are untouched
    # oracle on sampled users of the slice (host: the graph as scipy CSR)
    data = (G["y_val"] if weighted else torch.ones(G["nnz"], dtype=torch.float64, device=dev)).cpu().numpy()
    Ysp = sp.csr_matrix((data, G["y_idx"].cpu().numpy(), G["y_ptr"].cpu().numpy()), shape=(ns, nt))
    rng = np.random.default_rng(3)
    deg = np.diff(Ysp.indptr[s0:s0 + users + 1])
    cand = np.flatnonzero(deg <= 200)  # the oracle needs the rows of U these users reach: keep it in host memory
    rows = np.sort(rng.choice(cand, size=48 if exact else 12, replace=False)) + s0
    F, _ = o.two_layer_scores_sparse(Ysp, rows)
    got_i, got_v = i1.cpu().numpy(), v1.cpu().numpy()
    for j, r in enumerate(rows):
        dense = np.zeros(nt)
        sl = slice(F.indptr[j], F.indptr[j + 1])
        dense[F.indices[sl]] = F.data[sl]
        order = o.sortperm_rev(dense)[:topl]
        if exact:
            assert np.array_equal(got_i[r - s0], order), (degrees, weighted, int(r))
            assert np.array_equal(got_v[r - s0].view(np.uint64), dense[order].view(np.uint64)), (degrees, weighted, int(r))
        else:
            assert np.allclose(got_v[r - s0], dense[order], rtol=1e-12, atol=0.0), (degrees, int(r))
            assert np.allclose(dense[got_i[r - s0]], dense[order], rtol=1e-12, atol=0.0), (degrees, int(r))
