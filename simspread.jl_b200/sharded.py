"""Multi-GPU form of predict for query rows (one process per GPU).

Sharding of R = Xq * T,  T = (Xs' * (Y ./ ks)) ./ kf:
  * rank r owns the query-row slab Xq[r] / R[r]              -> no communication;
  * T is built cooperatively by target-column block: rank r holds Y[:, r], computes
    Wst[:, r] = Y[:, r] ./ ks and T[:, r] = (Xs' * Wst[:, r]) ./ kf with the replicated Xs; column
    blocks of the column-major T are contiguous, so ONE all-gather assembles T on every rank;
  * degrees: kf from the replicated Xs; ks = nnz_row(Xs) + sum_r nnz_row(Y[:, r]) -> all-reduce of
    an int32 vector (rank 0 contributes the Xs term); kt[r] = nnz_col(Y[:, r]) -> all-gather
    (needed only by the fused clean!).

The product path is inside the library: `Comm` / `ShardedQuery` below are thin callers of `ss_comm_*`,
`ss_sharded_*` and `ss_predict_query_sharded` (csrc/ss_comm.cu: NCCL + fused GEMM / all-gather over CUDA-IPC peer
mappings).  `ShardedPredict` restates the same choreography step by step over a backend object; it exists so that the
order of the exchange steps can be exercised on CPU: tests/test_sharded_gloo.py drives it at world_size 2 with a NumPy
double and the gloo backend."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass


@dataclass
class ShardPlan:
    world: int
    rank: int
    nq: int
    nt: int
    nq_blk: int  # rows per rank (ceil); the last ranks may own fewer (or zero) real rows
    nt_blk: int  # target columns per rank (ceil)

    @property
    def q0(self) -> int:
        return min(self.nq, self.rank * self.nq_blk)

    @property
    def nq_local(self) -> int:
        return max(0, min(self.nq_blk, self.nq - self.rank * self.nq_blk))

    @property
    def t0(self) -> int:
        return min(self.nt, self.rank * self.nt_blk)

    @property
    def nt_local(self) -> int:
        return max(0, min(self.nt_blk, self.nt - self.rank * self.nt_blk))

    @property
    def nt_padded(self) -> int:
        return self.nt_blk * self.world


def make_plan(nq: int, nt: int, world: int, rank: int) -> ShardPlan:
    assert world >= 1 and 0 <= rank < world
    return ShardPlan(world, rank, nq, nt, -(-nq // world), -(-nt // world))


class ShardedPredict:
    """Choreography of one sharded spread+predict step.  `backend` provides:
         degrees(with_xs_rows: bool)   fill ks_part (nnz_row(Y blk) [+ nnz_row(Xs)]), kf, kt_blk
         all_reduce_ks(), all_gather_kt(), all_gather_T()
         spread()                      Wst blk = Y blk ./ ks
         gemm_T()                      T blk = (Xs' * Wst blk) ./ kf
         gemm_R(clean: bool)           R slab = Xq slab * T   (+ clean! flag from kt)
    """

    def __init__(self, plan: ShardPlan, backend):
        self.plan, self.b = plan, backend

    def step(self, clean: bool = True):
        self.front()
        self.b.gemm_R(clean)

    def front(self):
        """Everything up to the assembled T (and kt) on every rank."""
        b = self.b
        b.degrees(with_xs_rows=(self.plan.rank == 0))
        if self.plan.world > 1:
            b.all_reduce_ks()
            b.all_gather_kt()
        b.spread()
        b.gemm_T()
        if self.plan.world > 1:
            b.all_gather_T()


class _CudaView:
    """Minimal __cuda_array_interface__ carrier: a torch view of library-owned device memory."""

    def __init__(self, ptr: int, shape, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class Comm:
    """One rank of the library's NCCL communicator (`ss_comm`, include/simspread_b200.h section 3b).  The collectives
    live behind the C ABI; this class only carries the unique id to the ranks: `exchange` is any callable that
    returns rank 0's 128 bytes on every rank (e.g. a torch.distributed / MPI broadcast), or pass `path` for the
    file rendezvous (`ss_comm_init_file`), which needs no other channel at all."""

    def __init__(self, ctx, rank: int, world: int, exchange=None, path: str = None, timeout_s: float = 120.0):
        from ._lib import check, lib
        self.ctx, self.rank, self.world, self.L, self.check = ctx, int(rank), int(world), lib(), check
        h = C.c_void_p()
        if path is not None:
            check(self.L.ss_comm_init_file(ctx.h, self.rank, self.world, path.encode(), float(timeout_s), C.byref(h)))
        else:
            buf = (C.c_ubyte * 128)()
            if self.world > 1:
                if self.rank == 0:
                    check(self.L.ss_comm_unique_id(buf))
                raw = exchange(bytes(buf))
                assert len(raw) == 128
                buf = (C.c_ubyte * 128).from_buffer_copy(raw)
            check(self.L.ss_comm_init(ctx.h, self.rank, self.world, buf, C.byref(h)))
        self.h = h

    def barrier(self):
        self.check(self.L.ss_comm_barrier(self.h))

    def allreduce_host(self, values, op: str = "sum"):
        """sum / max over the ranks of a short list of host doubles (timings, checksums)."""
        arr = (C.c_double * len(values))(*[float(v) for v in values])
        self.check(self.L.ss_comm_allreduce_host_f64(self.h, arr, len(values), 1 if op == "max" else 0))
        return [float(x) for x in arr]

    def nccl_version(self) -> int:
        v = C.c_int32()
        self.check(self.L.ss_comm_info(self.h, None, None, C.byref(v)))
        return int(v.value)

    def close(self):
        if getattr(self, "h", None):
            self.L.ss_comm_destroy(self.h)
            self.h = None

    __del__ = close


class ShardedQuery:
    """Thin caller of the sharded predict of the C ABI (`ss_sharded_*`, `ss_predict_query_sharded`): rank r passes its
    query-row slab Xq[r] (nq_local x nf), the replicated Xs (ns x nf), its target-column block Y[:, r] (ns x nt_blk)
    and receives R[r] (nq_local x nt).  Degrees, the two NCCL collectives and the T tiles stored into every rank's T
    from the GEMM epilogue all happen inside the library."""

    def __init__(self, comm: Comm, ns: int, nf: int, nt: int):
        self.comm, self.L, self.check = comm, comm.L, comm.check
        h = C.c_void_p()
        self.check(self.L.ss_sharded_create(comm.h, int(ns), int(nf), int(nt), C.byref(h)))
        self.h = h
        blk, fused = C.c_int64(), C.c_int32()
        self.check(self.L.ss_sharded_info(h, C.byref(blk), C.byref(fused)))
        self.nt_blk, self.fused = int(blk.value), bool(fused.value)

    def front(self, Xs, Yblk):
        self.check(self.L.ss_sharded_front(self.h, Xs.h, Yblk.h))

    def views(self):
        """(T, kt) handles owned by the plan (non-owning for the caller)."""
        t, k = C.c_void_p(), C.c_void_p()
        self.check(self.L.ss_sharded_views(self.h, C.byref(t), C.byref(k)))
        return t, k

    def predict(self, Xq, Xs, Yblk, R, clean: bool = True):
        from ._lib import SS_PREDICT_CLEAN
        self.check(self.L.ss_predict_query_sharded(self.h, Xq.h if Xq is not None else None, Xs.h, Yblk.h,
                                                   R.h if R is not None else None, SS_PREDICT_CLEAN if clean else 0))

    def close(self):
        if getattr(self, "h", None):
            self.L.ss_sharded_destroy(self.h)  # collective: closes the peer mappings after a barrier
            self.h = None


def combine_segment_summaries(sizes, summaries, rank: int):
    """global6 of `ss_auc_segment_integrate` for segment `rank`, from the (pairs, (positives, last run start,
    positives before it)) of every segment in ascending key order.  Pure host arithmetic (CPU-testable)."""
    P = sum(int(su[0]) for su in summaries)
    Mtot = sum(int(m) for m in sizes)
    idx0 = pos_below = 0
    init_start, init_startpos = -1, 0
    for r in range(rank):
        m, (pos, start, startpos) = int(sizes[r]), (int(x) for x in summaries[r])
        if start >= 0:
            init_start, init_startpos = idx0 + start, pos_below + startpos
        idx0 += m
        pos_below += pos
    return [P, Mtot, idx0, pos_below, init_start, init_startpos]


def pick_splitters(samples_u64, world: int):
    """world - 1 key splitters from the gathered, regularly spaced samples of every rank's sorted keys."""
    import numpy as np
    sm = np.sort(np.asarray(samples_u64, dtype=np.uint64).ravel())
    if sm.size == 0:
        return np.zeros(world - 1, dtype=np.uint64)
    return sm[[min(sm.size - 1, (i * sm.size) // world) for i in range(1, world)]].astype(np.uint64)


class LibAucBackend:
    """Device side of the sample-sort AuROC: libsimspread_b200 on this rank's GPU.  The sorted (key, label) arrays
    live in the context's sort buffers; `keys` / `labs` are torch views of them (uint64 keys moved as int64 bits)."""

    def __init__(self, ss, ctx, torch, dev):
        from ._lib import check
        self.L, self.ctx, self.torch, self.dev, self.check = ss.lib(), ctx, torch, dev, check
        self.pk, self.pl, self.m = C.c_void_p(), C.c_void_p(), 0

    def sort(self, labels, scores=None, keys=None):
        t = self.torch
        self.m = int(labels.numel())
        t.cuda.synchronize(self.dev)
        self.check(self.L.ss_auc_sort(self.ctx.h, C.c_void_p(labels.data_ptr()),
                                      C.c_void_p(scores.data_ptr()) if scores is not None else None,
                                      C.c_void_p(keys.data_ptr()) if keys is not None else None, self.m,
                                      C.byref(self.pk), C.byref(self.pl)))
        n = max(self.m, 1)
        return (t.as_tensor(_CudaView(self.pk.value, (n,), "<i8"), device=self.dev)[:self.m],
                t.as_tensor(_CudaView(self.pl.value, (n,), "|u1"), device=self.dev)[:self.m])

    def lower_bound(self, split_u64):
        import numpy as np
        cut = np.zeros(len(split_u64), dtype=np.int64)
        self.check(self.L.ss_auc_lower_bound(self.ctx.h, self.pk, self.m, split_u64.ctypes.data, len(split_u64), cut.ctypes.data))
        return cut

    def summary(self):
        import numpy as np
        out = np.zeros(3, dtype=np.int64)
        self.check(self.L.ss_auc_segment_summary(self.ctx.h, self.pk, self.pl, self.m, out.ctypes.data))
        return out

    def integrate(self, g6):
        import numpy as np
        g = np.asarray(g6, dtype=np.int64)
        part = (C.c_double * 2)()
        self.check(self.L.ss_auc_segment_integrate(self.ctx.h, self.pk, self.pl, self.m, g.ctypes.data, part))
        return float(part[0]), float(part[1])


def samplesort_auroc_auprc(backend, torch, dist, labels, scores, world: int, rank: int, samples_per_rank: int = 64):
    """Choreography of the distributed AuROC / AuPRC (see `global_auroc_auprc`).  `backend` does the per-rank work
    (LibAucBackend on a GPU; tests/test_sharded_gloo.py drives it with a NumPy double over gloo)."""
    import numpy as np
    dev = labels.device
    m = int(labels.numel())
    keys, labs = backend.sort(labels, scores=scores)
    # 1. splitters from regularly spaced samples of every rank's sorted keys
    ns = samples_per_rank
    samp = torch.zeros(ns, dtype=torch.int64, device=dev)
    have = torch.tensor([min(ns, m)], dtype=torch.int64, device=dev)
    if m:
        cnt = min(ns, m)  # integer arithmetic: a float32 linspace rounds past m - 1 for large m
        pos = (torch.arange(cnt, dtype=torch.int64, device=dev) * (m - 1)) // max(cnt - 1, 1)
        samp[:cnt] = keys[pos]
    all_s = [torch.zeros_like(samp) for _ in range(world)]
    all_n = [torch.zeros_like(have) for _ in range(world)]
    dist.all_gather(all_s, samp)
    dist.all_gather(all_n, have)
    gathered = np.concatenate([t.cpu().numpy().view(np.uint64)[:int(n.item())] for t, n in zip(all_s, all_n)])
    split = pick_splitters(gathered, world)
    # 2. every pair goes to the rank that owns its key range (keys < splitter[r] -> ranks <= r): one all-to-all
    cut = np.maximum.accumulate(backend.lower_bound(split)) if world > 1 else np.zeros(0, dtype=np.int64)
    send = np.diff(np.concatenate([[0], cut, [m]])).astype(np.int64)
    send_t = torch.from_numpy(send).to(dev)
    recv_t = torch.zeros_like(send_t)
    dist.all_to_all_single(recv_t, send_t)
    recv = recv_t.cpu().numpy()
    mr = int(recv.sum())
    rk = torch.empty(max(mr, 1), dtype=torch.int64, device=dev)[:mr]
    rl = torch.empty(max(mr, 1), dtype=torch.uint8, device=dev)[:mr]
    dist.all_to_all_single(rk, keys.contiguous(), output_split_sizes=recv.tolist(), input_split_sizes=send.tolist())
    dist.all_to_all_single(rl, labs.contiguous(), output_split_sizes=recv.tolist(), input_split_sizes=send.tolist())
    # 3. sort the received runs, exchange the 3-integer summaries, integrate this key range
    backend.sort(rl, keys=rk)
    summ = backend.summary()
    mine = torch.tensor([mr, int(summ[0]), int(summ[1]), int(summ[2])], dtype=torch.int64, device=dev)
    every = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    every = [t.cpu().numpy() for t in every]
    g6 = combine_segment_summaries([e[0] for e in every], [e[1:] for e in every], rank)
    part = backend.integrate(g6)
    out = torch.tensor([part[0], part[1]], dtype=torch.float64, device=dev)
    dist.all_reduce(out)
    return abs(float(out[0].item())), abs(float(out[1].item()))


def global_auroc_auprc(ss, ctx, torch, dist, y_local, r_local, world: int, rank: int, method: str = "samplesort"):
    """AuROC / AuPRC over the scores of ALL ranks (reference src/performance.jl:49-63, 74-89 on the
    concatenation of the per-rank (label, score) lists -- both areas are order-independent).

    method="samplesort" (default): every rank radix-sorts its pairs on the device (`ss_auc_sort`), the ranks agree
    on world-1 key splitters from regularly spaced samples, ONE NCCL all-to-all moves every pair to the rank that
    owns its key range (equal keys never straddle two ranks), each rank re-sorts what it received and integrates
    the trapezoids of its range given the counts below it (`ss_auc_segment_summary` / `_integrate`; the summaries
    are all-gathered, 3 integers per rank), and the signed partial areas are all-reduced.  Memory and work per
    rank stay ~1/world of the list.
    method="gather": NCCL gather of the slabs to rank 0 and the single-GPU kernel there (needs 27 B per score of
    the WHOLE list on rank 0)."""
    dev = r_local.device
    scores = r_local.reshape(-1).to(torch.float64).contiguous()
    labels = (y_local.reshape(-1) != 0).to(torch.uint8).contiguous()
    assert scores.numel() == labels.numel(), "The number of scores must be equal to the number of labels"
    if method == "gather" or world == 1:
        return _gather_auroc_auprc(ss, ctx, torch, dist, labels, scores, world, rank)
    assert method == "samplesort"
    return samplesort_auroc_auprc(LibAucBackend(ss, ctx, torch, dev), torch, dist, labels, scores, world, rank)


def _gather_auroc_auprc(ss, ctx, torch, dist, labels, scores, world: int, rank: int):
    from ._lib import check
    dev = scores.device
    n = torch.tensor([scores.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    if world > 1:
        dist.all_gather(sizes, n)
    else:
        sizes = [n]
    sizes = [int(s.item()) for s in sizes]
    out = torch.zeros(2, dtype=torch.float64, device=dev)
    if world == 1:
        all_s, all_l = scores, labels
    else:
        cap = max(sizes)
        ps = torch.zeros(cap, dtype=torch.float64, device=dev)
        pl = torch.zeros(cap, dtype=torch.uint8, device=dev)
        ps[:scores.numel()] = scores
        pl[:labels.numel()] = labels
        gs = [torch.empty(cap, dtype=torch.float64, device=dev) for _ in range(world)] if rank == 0 else None
        gl = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(ps, gs, dst=0)
        dist.gather(pl, gl, dst=0)
        if rank == 0:
            all_s = torch.cat([g[:m] for g, m in zip(gs, sizes)])
            all_l = torch.cat([g[:m] for g, m in zip(gl, sizes)])
    if rank == 0:
        res = (C.c_double * 2)()
        torch.cuda.synchronize(dev)
        check(ss.lib().ss_auroc_auprc(ctx.h, C.c_void_p(all_l.data_ptr()), C.c_void_p(all_s.data_ptr()),
                                      int(all_s.numel()), res))
        out[0], out[1] = float(res[0]), float(res[1])
    if world > 1:
        dist.broadcast(out, src=0)
    return float(out[0].item()), float(out[1].item())
