"""CPU replay of the edge-list form of the first product (simspread.jl_b200/csrc/ss_tsparse.cu; no GPU needed):
W = Y ./ ks compacted by target column (sources ascending; Inf / NaN weights -> 0 as spread() does, src/core.jl:365-371),
and the walk of `tsp_kernel` over slabs of 64 sources with 32-edge windows per target: the edges of a slab are a prefix
of the window (counted with a ballot), a full window inside one slab is followed by the next one, cursors carry over to
the next slab.  Every T[f,t] must come out as the sum over ALL edges of the column in ascending source order, divided by
kf (0 where kf == 0) -- i.e. equal to the dense form up to rounding and independent of how the targets are tiled."""
import numpy as np
import pytest

SB, WIN = 64, 32


def wcsc(Y, ks):
    col_ptr, row_idx, val = [0], [], []
    for t in range(Y.shape[1]):
        for s in range(Y.shape[0]):
            v = Y[s, t]
            if v != 0.0:                      # count(!iszero): NaN is an edge
                with np.errstate(invalid="ignore", divide="ignore"):
                    w = v / np.float64(ks[s])
                row_idx.append(s)
                val.append(w if np.isfinite(w) else 0.0)
        col_ptr.append(len(row_idx))
    return np.array(col_ptr), np.array(row_idx, dtype=np.int64), np.array(val)


def tsp_walk(Xs, col_ptr, row_idx, val, kf, targets):
    ns, nf = Xs.shape
    T = np.zeros((nf, len(targets)))
    visits = 0
    for jt, t in enumerate(targets):
        cur, end = int(col_ptr[t]), int(col_ptr[t + 1])
        acc = np.zeros(nf)
        for slab in range(-(-ns // SB)):
            s_hi = (slab + 1) * SB
            while True:
                win = [int(row_idx[c]) if c < end else 2**31 - 1 for c in range(cur, cur + WIN)]
                n = sum(1 for s in win if s < s_hi)
                assert all(s < s_hi for s in win[:n]) and all(s >= s_hi for s in win[n:])  # a prefix (sources ascend)
                for k in range(n):
                    s = win[k]
                    assert slab * SB <= s < s_hi                                            # inside the staged slab
                    acc = acc + val[cur + k] * Xs[s, :]
                    visits += 1
                cur += n
                if n < WIN:
                    break
        assert cur == end
        with np.errstate(invalid="ignore", divide="ignore"):
            T[:, jt] = np.where(kf > 0, acc / np.maximum(kf, 1), 0.0)
    return T, visits


@pytest.mark.parametrize("ns,nf,nt,dens", [(1, 1, 1, 1.0), (130, 7, 9, 0.05), (200, 5, 6, 0.7), (64, 3, 4, 1.0), (129, 4, 3, 0.0)])
def test_edge_list_walk_equals_the_dense_form(ns, nf, nt, dens):
    rng = np.random.default_rng(ns + nt)
    Xs = np.where(rng.random((ns, nf)) < 0.6, np.round(rng.random((ns, nf)), 3), 0.0)
    Y = (rng.random((ns, nt)) < dens).astype(float)
    if ns > 10 and nt > 2:
        Y[3, 1], Y[5, 2] = np.nan, np.inf           # edges whose weight spread() sets to 0
    ks = (Xs != 0).sum(1) + (Y != 0).sum(1)
    kf = (Xs != 0).sum(0)
    col_ptr, row_idx, val = wcsc(Y, ks)
    assert np.array_equal(np.diff(col_ptr), (Y != 0).sum(0))                                # == the degree kernel's kt
    T, visits = tsp_walk(Xs, col_ptr, row_idx, val, kf, list(range(nt)))
    assert visits == int((Y != 0).sum())
    with np.errstate(invalid="ignore", divide="ignore"):
        W = Y / np.maximum(ks, 1)[:, None]
    W[~np.isfinite(W)] = 0.0
    want = np.where(kf[:, None] > 0, (Xs.T @ W) / np.maximum(kf, 1)[:, None], 0.0)
    assert np.allclose(T, want, rtol=1e-13, atol=0.0)
    # tiling of the targets (CTA tiles, shards of a multi-GPU run) does not change a single bit
    order = list(rng.permutation(nt))
    T2, _ = tsp_walk(Xs, col_ptr, row_idx, val, kf, order)
    assert np.array_equal(T2, T[:, order])
