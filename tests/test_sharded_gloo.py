"""World-size-2 test (gloo, CPU) of the multi-GPU choreography in simspread.jl_b200/sharded.py:
partition plan, all-reduce of ks, all-gather of kt and of the T column blocks.  The numerical
kernels are replaced by a NumPy test double built from the oracle -- this checks the *plumbing* of
the N > 1 path; the kernels themselves are checked by the `-m gpu` tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class NumpyBackend:
    """Test double with the LibBackend interface (oracle arithmetic, gloo collectives)."""

    def __init__(self, o, plan, Xq, Xs, Yblk):
        self.o, self.p = o, plan
        self.Xq, self.Xs, self.Y = Xq, Xs, Yblk  # Y block padded to nt_blk columns
        self.R = None

    def degrees(self, with_xs_rows):
        nzx, nzy = self.Xs != 0, self.Y != 0
        self.ks = nzy.sum(1).astype(np.int32) + (nzx.sum(1).astype(np.int32) if with_xs_rows else 0)
        self.kf = nzx.sum(0).astype(np.int32)
        self.ktl = nzy.sum(0).astype(np.int32)
        self.kt = self.ktl

    def all_reduce_ks(self):
        t = torch.from_numpy(np.ascontiguousarray(self.ks))
        dist.all_reduce(t)
        self.ks = t.numpy()

    def all_gather_kt(self):
        out = torch.zeros(self.p.nt_padded, dtype=torch.int32)
        dist.all_gather_into_tensor(out, torch.from_numpy(np.ascontiguousarray(self.ktl)))
        self.kt = out.numpy()

    def spread(self):
        self.W = self.o._div_rows(self.Y, self.ks)

    def gemm_T(self):
        self.Tl = self.o._div_rows(np.ascontiguousarray(self.Xs.T) @ self.W, self.kf)  # (nf, nt_blk)
        self.T = self.Tl

    def all_gather_T(self):
        # column-major T: a column block is contiguous -> gather the transposed (row-major) blocks
        out = torch.zeros((self.p.nt_padded, self.Tl.shape[0]), dtype=torch.float64)
        dist.all_gather_into_tensor(out, torch.from_numpy(np.ascontiguousarray(self.Tl.T)))
        self.T = out.numpy().T

    def gemm_R(self, clean):
        nt = self.p.nt
        self.R = self.Xq @ self.T[:, :nt]
        if clean:
            self.R[:, self.kt[:nt] == 0] = -99.0


def _worker(rank, world, port, nq, ns, nf, nt, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import simspread_b200  # noqa: F401  (loads the package so that the submodule import below works)
    from simspread_b200.sharded import ShardedPredict, make_plan
    from oracle import simspread_oracle as o
    Xq, Xs, Y = o.synth_dense(nq, ns, nf, nt, seed=3, y_density=0.2, alpha=0.3, weighted=True)
    Y[:, 1] = 0.0  # a target without edges -> clean! flag must survive the all-gather of kt
    plan = make_plan(nq, nt, world, rank)
    Yblk = np.zeros((ns, plan.nt_blk))
    Yblk[:, :plan.nt_local] = Y[:, plan.t0:plan.t0 + plan.nt_local]
    be = NumpyBackend(o, plan, Xq[plan.q0:plan.q0 + plan.nq_local], Xs, Yblk)
    ShardedPredict(plan, be).step(clean=True)
    want = o.predict_blocks_query(Xq, Xs, Y)
    o.clean_blocks(want, o.degrees_blocks(Xs, Y)[2])
    mine = want[plan.q0:plan.q0 + plan.nq_local]
    ok = be.R.shape == mine.shape and np.allclose(be.R, mine, rtol=1e-13, atol=0) and \
        np.array_equal(be.R == -99, mine == -99)
    ks, kf, kt = o.degrees_blocks(Xs, Y)
    ok = ok and np.array_equal(be.ks, ks) and np.array_equal(be.kt[:nt], kt)
    q.put((rank, bool(ok), plan.q0, plan.nq_local, plan.t0, plan.nt_local))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nq,ns,nf,nt", [(8, 12, 10, 6), (7, 9, 11, 5)])
def test_sharded_choreography_world2(nq, ns, nf, nt):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nq, ns, nf, nt, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    # the slabs tile the query rows / target columns exactly once
    assert sum(r[3] for r in res) == nq and sum(r[5] for r in res) == nt
    assert [r[2] for r in res] == [0, -(-nq // 2)]


def test_plan_partitions():
    sys.path.insert(0, ROOT)
    import simspread_b200  # noqa: F401
    from simspread_b200.sharded import make_plan
    for n, w in [(100000, 8), (50000, 8), (7, 2), (5, 4), (3, 8)]:
        plans = [make_plan(n, n, w, r) for r in range(w)]
        assert sum(p.nq_local for p in plans) == n
        cover = []
        for p in plans:
            cover += list(range(p.q0, p.q0 + p.nq_local))
        assert cover == list(range(n))
        assert all(p.nt_padded >= n and p.nt_padded - n < w for p in plans)


# ---------------------------------------------------------------------------------------------
# AuROC / AuPRC of a list spread over the ranks (sample sort + per-range integration), SURVEY 8f-1
# ---------------------------------------------------------------------------------------------


class NumpyAucBackend:
    """Test double with the LibAucBackend interface: the per-rank device work restated in NumPy
    (keys = the oracle's isless image of the scores, moved around as int64 bit patterns)."""

    def __init__(self, o):
        self.o = o
        self.k = np.zeros(0, np.uint64)
        self.l = np.zeros(0, np.uint8)

    def sort(self, labels, scores=None, keys=None):
        lab = labels.numpy().astype(np.uint8)
        k = self.o._isless_key(scores.numpy()).astype(np.uint64) if scores is not None else keys.numpy().view(np.uint64)
        order = np.argsort(k, kind="stable")
        self.k, self.l = k[order], lab[order]
        return torch.from_numpy(self.k.view(np.int64).copy()), torch.from_numpy(self.l.copy())

    def lower_bound(self, split):
        return np.searchsorted(self.k, split, side="left").astype(np.int64)

    def _starts(self):
        if self.k.size == 0:
            return np.zeros(0, np.int64)
        return np.flatnonzero(np.concatenate([[True], self.k[1:] != self.k[:-1]])).astype(np.int64)

    def summary(self):
        st = self._starts()
        if st.size == 0:
            return np.array([0, -1, 0], dtype=np.int64)
        return np.array([int(self.l.sum()), int(st[-1]), int(self.l[:st[-1]].sum())], dtype=np.int64)

    def integrate(self, g6):
        P, Mtot, idx0, pos_below, init_start, init_startpos = (int(x) for x in g6)
        N = Mtot - P
        st = self._starts()
        cum = np.concatenate([[0], np.cumsum(self.l.astype(np.int64))])
        pts = [(idx0 + int(s), pos_below + int(cum[s])) for s in st]  # (global index of a run start, positives before it)
        if init_start >= 0:
            pts = [(init_start, init_startpos)] + pts
        roc = pr = 0.0
        for (p1, b1), (p2, b2) in zip(pts[:-1], pts[1:]):  # lower threshold, higher threshold
            tp1, fp1, tp2, fp2 = P - b1, N - (p1 - b1), P - b2, N - (p2 - b2)
            roc += (fp2 / N - fp1 / N) * (tp1 / P + tp2 / P) * 0.5
            pr += (tp2 / P - tp1 / P) * (tp1 / (tp1 + fp1) + tp2 / (tp2 + fp2)) * 0.5
        return roc, pr


def _auc_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import simspread_b200  # noqa: F401
    from simspread_b200.sharded import samplesort_auroc_auprc
    from oracle import simspread_oracle as o
    ok = True
    for case, sizes in enumerate([(700, 1300), (0, 900), (5, 3)]):
        rng = np.random.default_rng(100 + case)
        parts = []
        for m in sizes:  # every rank builds all slabs, keeps its own: ties inside and across slabs, -0.0 / 0.0
            sc = np.round(rng.random(m), 2)
            sc[: m // 10] = 0.0
            sc[m // 10: m // 8] = -0.0
            lb = (rng.random(m) < 0.2 + 0.5 * sc).astype(np.uint8)
            parts.append((sc, lb))
        if sum(int(p[1].sum()) for p in parts) == 0:
            parts[-1][1][0] = 1
        sc, lb = parts[rank]
        got = samplesort_auroc_auprc(NumpyAucBackend(o), torch, dist, torch.from_numpy(lb.copy()), torch.from_numpy(sc.copy()),
                                     world, rank, samples_per_rank=8)
        S = np.concatenate([p[0] for p in parts])
        Lb = np.concatenate([p[1] for p in parts])
        want = (o.AuROC(Lb > 0, S), o.AuPRC(Lb > 0, S))
        ok = ok and abs(got[0] - want[0]) <= 1e-12 * abs(want[0]) and abs(got[1] - want[1]) <= 1e-12 * abs(want[1])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_samplesort_auroc_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_auc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


def test_segment_summaries_combine():
    sys.path.insert(0, ROOT)
    import simspread_b200  # noqa: F401
    from simspread_b200.sharded import combine_segment_summaries, pick_splitters
    sizes = [4, 0, 3, 5]
    summ = [(2, 3, 1), (0, -1, 0), (1, 0, 0), (3, 2, 1)]
    assert combine_segment_summaries(sizes, summ, 0) == [6, 12, 0, 0, -1, 0]
    assert combine_segment_summaries(sizes, summ, 1) == [6, 12, 4, 2, 3, 1]
    assert combine_segment_summaries(sizes, summ, 2) == [6, 12, 4, 2, 3, 1]       # the empty segment changes nothing
    assert combine_segment_summaries(sizes, summ, 3) == [6, 12, 7, 3, 4, 2]       # last run start below: 4 + 0
    sp = pick_splitters(np.array([5, 1, 9, 3, 7, 2, 8, 4], dtype=np.uint64), 4)
    assert sp.dtype == np.uint64 and list(sp) == [3, 5, 8] and len(pick_splitters(np.zeros(0, np.uint64), 3)) == 2
