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
