"""`-m gpu` parity tests of the sparse recommender form (BASELINE config 5): `predict(A, ytrain)` of the reference
(src/core.jl:446-466) on the 2-layer graph [0 Y; Y' 0], reduced to `sortperm(rev=true)[1:L]` per source
(src/performance.jl:315) by `ss_recommend_topl`.

The bar for this path is the north star's "bit-exact top-k order".  The kernels evaluate the reference's association
A * (W * W) with every sum in ascending index order (ss_transfer.cu), which is also what a CSR x CSR product of
scipy.sparse does -- so scores, and with them the order of tied scores, must equal the oracle BIT FOR BIT, on binary
graphs (where exact ties are everywhere) as well as on weighted ones, and two runs must agree bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ss():
    import simspread_b200 as m
    m.build()
    m.Context.default()
    return m


@pytest.fixture(scope="module")
def o():
    from oracle import simspread_oracle
    return simspread_oracle


def graph(rng, users, items, dens, weighted, degrees="uniform"):
    if degrees == "pareto":  # heavy-tailed user activity, a few very popular items
        pu = np.minimum(1.0, dens * 0.3 * rng.random(users) ** -0.7)[:, None]
        pi = np.minimum(1.0, 0.5 + 8.0 * (rng.random(items) < 0.01))[None, :]
        mask = rng.random((users, items)) < np.minimum(1.0, pu * pi)
    else:
        mask = rng.random((users, items)) < dens
    Y = np.where(mask, np.round(rng.random((users, items)) + 0.5, 3), 0.0) if weighted else mask.astype(float)
    return Y


def dense_scores(o, Y):
    F, U = o.two_layer_scores_sparse(Y)
    return F.toarray(), U


CASES = [(300, 200, 0.05, 16, "uniform"), (1000, 2600, 0.01, 16, "uniform"), (64, 40000, 0.002, 32, "uniform"),
         (40, 3000, 0.5, 20, "uniform"), (3, 70001, 0.01, 8, "uniform"), (500, 13000, 0.004, 20, "pareto"),
         (200, 6144, 0.02, 20, "uniform"), (150, 6145, 0.03, 1, "uniform"), (20, 1537, 0.3, 32, "uniform")]


@pytest.mark.parametrize("users,items,dens,L,degrees", CASES)
@pytest.mark.parametrize("weighted", [False, True])
def test_recommend_topl_is_order_exact_and_deterministic(ss, o, users, items, dens, L, degrees, weighted):
    """idx == sortperm_rev(F)[:L] row for row and val == F[idx] bit for bit against the oracle (scipy.sparse sums in the
    same order), on binary and weighted graphs; a second run returns identical bits."""
    rng = np.random.default_rng(users * 7 + items + int(weighted))
    Y = graph(rng, users, items, dens, weighted, degrees)
    Y[min(3, users - 1), :] = 0.0       # a user without items: every score 0 -> the first L columns
    if items > 10:
        Y[:, 7] = 0.0                   # an item nobody has
    F, _ = dense_scores(o, Y)
    idx, val = ss.recommend_topl(Y, L, weighted=weighted)
    order = np.stack([o.sortperm_rev(F[u])[:L] for u in range(users)])
    assert np.array_equal(idx, order)
    assert np.array_equal(val, np.take_along_axis(F, order, axis=1))
    idx2, val2 = ss.recommend_topl(Y, L, weighted=weighted)
    assert np.array_equal(idx, idx2) and np.array_equal(val.view(np.uint64), val2.view(np.uint64))
    assert np.array_equal(idx[min(3, users - 1)], np.arange(L))
    # cross-check against the dense block form (BLAS order): same scores within the FP64 tolerance
    ks, kt = np.count_nonzero(Y, axis=1), np.count_nonzero(Y, axis=0)
    Fd = Y @ (o._div_rows(np.ascontiguousarray(Y.T), kt) @ o._div_rows(Y, ks))
    assert np.allclose(val, -np.sort(-Fd, axis=1)[:, :L], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("weighted", [False, True])
def test_transfer_matrix_equals_scipy_product(ss, o, weighted):
    """ss_transfer_build: U = (Y' ./ kt) * (Y ./ ks), rows sorted by column, bit-equal to the scipy.sparse product."""
    from simspread_b200._lib import check
    rng = np.random.default_rng(5 + int(weighted))
    users, items = 400, 13000   # 7 column tiles of 2048
    Y = graph(rng, users, items, 0.01, weighted)
    Y[:, :40] = np.where(rng.random((users, 40)) < 0.3, 1.0 if not weighted else 0.75, 0.0)  # popular items: many common raters
    _, U = dense_scores(o, Y)
    U.sort_indices()
    ctx = ss.Context.default()
    d = ss.DMat.from_host(ctx, Y)
    thr = 5e-324 if not weighted else float("-inf")
    cy = ss.DCsr.from_dense(ctx, d, thr, weighted)
    cyt = ss.DCsr.from_dense(ctx, d, thr, weighted, by_columns=True)
    h = C.c_void_p()
    check(ss.lib().ss_transfer_build(ctx.h, cy.h, cyt.h, C.byref(h)))
    try:
        info = (C.c_int64 * 4)()
        check(ss.lib().ss_transfer_info(h, info))
        assert info[0] == U.nnz and info[3] == -(-items // info[2]) and info[2] in (1024, 2048, 3072, 4096)
        rp = np.zeros(items + 1, np.int64)
        ci = np.zeros(max(1, info[0]), np.int32)
        va = np.zeros(max(1, info[0]), np.float64)
        check(ss.lib().ss_transfer_download(h, rp.ctypes.data, ci.ctypes.data, va.ctypes.data))
        assert np.array_equal(rp, U.indptr.astype(np.int64))
        assert np.array_equal(ci[:U.nnz], U.indices)
        assert np.array_equal(va[:U.nnz].view(np.uint64), U.data.view(np.uint64))
        # ranking a sub-range of the sources from the prebuilt U (the multi-GPU shard call)
        L = 20
        idx, val = ss.DIVec(ctx, L * users), ss.DMat(ctx, L, users)
        check(ss.lib().ss_recommend_topl_transfer(ctx.h, cy.h, h, L, 100, 250, idx.h, val.h))
        F, _ = dense_scores(o, Y)
        order = np.stack([o.sortperm_rev(F[u])[:L] for u in range(100, 250)])
        assert np.array_equal(idx.to_host().reshape(users, L)[100:250], order)
        assert np.array_equal(val.to_host().T[100:250], np.take_along_axis(F[100:250], order, axis=1))
    finally:
        check(ss.lib().ss_transfer_destroy(h))


def test_recommend_topl_tile_chunks_and_tile_ranges(ss, o, monkeypatch):
    """U that does not fit the memory budget is processed in chunks of column tiles with a running top-L; splitting a
    source over ranges of tiles (done when there are fewer sources than warps) gives the same bits."""
    rng = np.random.default_rng(77)
    users, items, L = 300, 40000, 20
    Y = graph(rng, users, items, 0.004, False)
    F, U = dense_scores(o, Y)
    order = np.stack([o.sortperm_rev(F[u])[:L] for u in range(users)])
    want = np.take_along_axis(F, order, axis=1)
    for parts in ("1", "3", "1000"):
        monkeypatch.setenv("SS_RECSYS_PARTS", parts)
        idx, val = ss.recommend_topl(Y, L, weighted=False)
        assert np.array_equal(idx, order) and np.array_equal(val, want), parts
    monkeypatch.delenv("SS_RECSYS_PARTS")
    # every tile shape of the streaming kernel gives the same bits (the default width is chosen from the graph)
    for tile in ("1024", "3072", "4096", "2048"):
        monkeypatch.setenv("SS_RECSYS_TILE", tile)
        idx, val = ss.recommend_topl(Y, L, weighted=False)
        assert np.array_equal(idx, order) and np.array_equal(val, want), tile
        Yw = np.where(Y != 0, 0.5 + (np.arange(items) % 7)[None, :] / 8.0, 0.0)
        Fw, _ = dense_scores(o, Yw)
        ow = np.stack([o.sortperm_rev(Fw[u])[:L] for u in range(users)])
        idx, val = ss.recommend_topl(Yw, L, weighted=True)
        assert np.array_equal(idx, ow) and np.array_equal(val, np.take_along_axis(Fw, ow, axis=1)), tile
    monkeypatch.delenv("SS_RECSYS_TILE")
    monkeypatch.setenv("SS_RECSYS_U_MB", "1")   # U is ~ 10 B x nnz(U): force several chunks
    assert U.nnz * 10 > 3 * (1 << 20)
    idx, val = ss.recommend_topl(Y, L, weighted=False)
    assert np.array_equal(idx, order) and np.array_equal(val, want)
    monkeypatch.delenv("SS_RECSYS_U_MB")
    # sub-range of sources (sharding): untouched rows keep their previous content
    idx2, val2 = ss.recommend_topl(Y, L, weighted=False, s_range=(37, 120))
    assert np.array_equal(idx2[37:120], order[37:120]) and np.array_equal(val2[37:120], want[37:120])


def test_recommend_topl_massive_ties(ss, o):
    """Rows whose scores tie massively (block-structured binary graph, isolated users): the tie order is by column."""
    users, items, L = 96, 7000, 32
    Y = np.zeros((users, items))
    for u in range(users):          # users in 4 groups that share exactly the same items -> identical scores
        g = u % 4
        Y[u, g * 1500:(g * 1500 + 600)] = 1.0
    Y[5, :] = 0.0
    Y[6, :] = 0.0
    Y[6, 6999] = 1.0                # a user whose only item nobody else has
    F, _ = dense_scores(o, Y)
    idx, val = ss.recommend_topl(Y, L, weighted=False)
    order = np.stack([o.sortperm_rev(F[u])[:L] for u in range(users)])
    assert np.array_equal(idx, order)
    assert np.array_equal(val, np.take_along_axis(F, order, axis=1))


def test_atomic_mode_still_available(ss, o, monkeypatch):
    """The round-1 cluster kernel (red.global.add.f64, unordered sums) stays selectable for A/B runs; it matches
    within the FP64 tolerance only."""
    rng = np.random.default_rng(3)
    Y = graph(rng, 200, 3000, 0.02, True)
    F, _ = dense_scores(o, Y)
    monkeypatch.setenv("SS_RECSYS_MODE", "atomic")
    idx, val = ss.recommend_topl(Y, 16, weighted=True)
    assert np.allclose(val, -np.sort(-F, axis=1)[:, :16], rtol=1e-12, atol=1e-300)
