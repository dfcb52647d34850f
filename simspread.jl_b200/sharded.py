"""Multi-GPU form of predict for query rows (one process per GPU).

Sharding of R = Xq * T,  T = (Xs' * (Y ./ ks)) ./ kf:
  * rank r owns the query-row slab Xq[r] / R[r]              -> no communication;
  * T is built cooperatively by target-column block: rank r holds Y[:, r], computes
    Wst[:, r] = Y[:, r] ./ ks and T[:, r] = (Xs' * Wst[:, r]) ./ kf with the replicated Xs; column
    blocks of the column-major T are contiguous, so ONE all-gather assembles T on every rank;
  * degrees: kf from the replicated Xs; ks = nnz_row(Xs) + sum_r nnz_row(Y[:, r]) -> all-reduce of
    an int32 vector (rank 0 contributes the Xs term); kt[r] = nnz_col(Y[:, r]) -> all-gather
    (needed only by the fused clean!).

`ShardedPredict.step()` is the choreography; the numerical work is delegated to a backend.  The
product backend (`LibBackend`) calls libsimspread_b200 on device buffers owned by torch and uses
torch.distributed (NCCL) for the three collectives.  tests/test_sharded_gloo.py drives the same
choreography at world_size 2 on CPU with a NumPy test double and the gloo backend."""
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

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


class LibBackend:
    """Product backend: libsimspread_b200 kernels on torch-owned device buffers + NCCL."""

    def __init__(self, ss, ctx, torch, dist, plan: ShardPlan, ns: int, nf: int, bXq, ldq, bXs, lds, bY, ldy, bR, ldr):
        from ._lib import check
        self.ss, self.ctx, self.torch, self.dist, self.plan, self.check = ss, ctx, torch, dist, plan, check
        self.L = ss.lib()
        dev = bXs.device
        p = plan
        self.ns, self.nf = ns, nf
        ld16 = lambda n: (n + 15) // 16 * 16
        self.mXq = ss.DMat.wrap(ctx, bXq.data_ptr(), p.nq_local, nf, ldq)
        self.mXs = ss.DMat.wrap(ctx, bXs.data_ptr(), ns, nf, lds)
        self.mY = ss.DMat.wrap(ctx, bY.data_ptr(), ns, p.nt_blk, ldy)
        self.mR = ss.DMat.wrap(ctx, bR.data_ptr(), p.nq_local, p.nt, ldr)
        self.ldt = ld16(nf)
        self.bW = torch.zeros((p.nt_blk, ld16(ns)), dtype=torch.float64, device=dev)
        self.mW = ss.DMat.wrap(ctx, self.bW.data_ptr(), ns, p.nt_blk, ld16(ns))
        self.tks = torch.zeros(ns, dtype=torch.int32, device=dev)
        self.tkf = torch.zeros(nf, dtype=torch.int32, device=dev)
        self.tkt = torch.zeros(p.nt_padded, dtype=torch.int32, device=dev)
        self.tktl = torch.zeros(p.nt_blk, dtype=torch.int32, device=dev)
        # T is library-owned (plain cudaMalloc) so that its IPC handle can be mapped by the peers;
        # bT is a torch view of the same memory (checks, NCCL fallback).
        self.mTfull = ss.DMat(ctx, nf, p.nt_padded, ipc=True)
        _, _, ldt_, pT = self.mTfull.info()
        assert ldt_ == self.ldt
        self.bT = torch.as_tensor(_CudaView(pT, (p.nt_padded, self.ldt)), device=dev)
        self.mT = ss.DMat.wrap(ctx, pT, nf, p.nt, self.ldt)
        blk_off = p.rank * p.nt_blk * self.ldt * 8           # my column block inside any rank's T
        self.mTl = ss.DMat.wrap(ctx, pT + blk_off, nf, p.nt_blk, self.ldt)
        self.bTl = self.bT[p.rank * p.nt_blk:(p.rank + 1) * p.nt_blk]
        self.mirrors = None
        import os
        if p.world > 1 and os.environ.get("SS_FUSED_ALLGATHER", "1") != "0":
            self._open_peers(pT, blk_off)
        self.vks, self.vkf = self._ivec(self.tks), self._ivec(self.tkf)
        self.vktl = self._ivec(self.tktl)
        self.vkt = self._ivec(self.tkt[:p.nt])

    def _open_peers(self, pT, blk_off):
        """Exchange CUDA IPC handles of T and map every peer's T: the T-GEMM epilogue then stores this
        rank's column block straight into all peers (fused GEMM + all-gather over NVLink)."""
        torch, dist, L = self.torch, self.dist, self.L
        hbuf = (C.c_ubyte * 64)()
        self.check(L.ss_mat_ipc_handle(self.ctx.h, self.mTfull.h, hbuf))
        mine = torch.tensor(list(hbuf), dtype=torch.uint8, device=self.bT.device)
        allh = torch.zeros(64 * self.plan.world, dtype=torch.uint8, device=self.bT.device)
        dist.all_gather_into_tensor(allh, mine)
        allh = allh.cpu().numpy().reshape(self.plan.world, 64)
        ptrs = []
        try:
            for r in range(self.plan.world):
                if r == self.plan.rank:
                    continue
                raw = (C.c_ubyte * 64)(*allh[r].tolist())
                dp = C.c_void_p()
                self.check(L.ss_ipc_open(self.ctx.h, raw, C.byref(dp)))
                ptrs.append(dp.value)
        except Exception as e:  # no P2P / IPC on this box: keep the NCCL all-gather
            print(f"[simspread_b200] rank {self.plan.rank}: peer mapping failed ({e}); using NCCL all-gather")
            ptrs = None
        ok = torch.tensor([1 if ptrs is not None else 0], device=self.bT.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            self.peer_bases = ptrs
            self.mirrors = (C.c_void_p * len(ptrs))(*[p_ + blk_off for p_ in ptrs])

    def _ivec(self, t):
        v = self.ss.DIVec.__new__(self.ss.DIVec)
        v.ctx, v.n = self.ctx, t.numel()
        h = C.c_void_p()
        self.check(self.L.ss_ivec_wrap(self.ctx.h, C.c_void_p(t.data_ptr()), t.numel(), C.byref(h)))
        v.h = h
        return v

    def _wait_collective(self):
        self.torch.cuda.current_stream().synchronize()

    def degrees(self, with_xs_rows: bool):
        L, c = self.L, self.ctx.h
        if with_xs_rows:
            self.check(L.ss_degrees(c, self.mXs.h, self.mY.h, self.vks.h, self.vkf.h, self.vktl.h))
        else:
            self.check(L.ss_degrees(c, self.mXs.h, self.mY.h, None, self.vkf.h, self.vktl.h))
            self.check(L.ss_k_rows(c, self.mY.h, self.vks.h))
        if self.plan.world == 1:
            self.tkt[:self.plan.nt_blk].copy_(self.tktl)
            self._wait_collective()

    def all_reduce_ks(self):
        self.dist.all_reduce(self.tks)
        self._wait_collective()

    def all_gather_kt(self):
        self.dist.all_gather_into_tensor(self.tkt, self.tktl)
        self._wait_collective()

    def spread(self):
        self.check(self.L.ss_spread_rows(self.ctx.h, self.mY.h, self.vks.h, self.mW.h))

    def gemm_T(self):
        from ._lib import SS_OP_T
        if self.mirrors is not None:
            self.check(self.L.ss_gemm_f64_mirrored(self.ctx.h, SS_OP_T, self.mXs.h, self.mW.h, self.mTl.h, self.vkf.h,
                                                   None, len(self.mirrors), self.mirrors))
        else:
            self.check(self.L.ss_gemm_f64(self.ctx.h, SS_OP_T, self.mXs.h, self.mW.h, self.mTl.h, self.vkf.h, None))

    def all_gather_T(self):
        if self.mirrors is not None:
            # every rank's epilogue has already written its block into every T (the GEMM call
            # returns after its stream has drained); a barrier orders those writes before the R GEMM
            self.dist.barrier()
        else:
            self.dist.all_gather_into_tensor(self.bT.view(-1), self.bTl.reshape(-1).clone())
        self._wait_collective()

    def gemm_R(self, clean: bool):
        from ._lib import SS_OP_N
        if self.plan.nq_local == 0:
            return
        self.check(self.L.ss_gemm_f64(self.ctx.h, SS_OP_N, self.mXq.h, self.mT.h, self.mR.h, None,
                                      self.vkt.h if clean else None))


def global_auroc_auprc(ss, ctx, torch, dist, y_local, r_local, world: int, rank: int):
    """AuROC / AuPRC over the scores of ALL ranks (reference src/performance.jl:49-63, 74-89 on the
    concatenation of the per-rank (label, score) lists -- both areas are order-independent).

    Exchange step: the per-rank score slabs (float64) and labels (uint8) are gathered on rank 0 with one NCCL
    gather each (padded to the largest slab), and rank 0 runs the device sort + scan (`ss_auroc_auprc`) on the
    concatenation; the result is broadcast.  Needs 9 B per score plus the 18 B per score of sort buffers on rank
    0 (C4 on 8 GPUs: 45 GB + 90 GB of 180 GB).  A sample-sort over the ranks would remove that limit; per-query
    metrics (recall@L / precision@L) need no exchange at all."""
    from ._lib import check
    dev = r_local.device
    scores = r_local.reshape(-1).to(torch.float64)
    labels = (y_local.reshape(-1) != 0).to(torch.uint8)
    assert scores.numel() == labels.numel(), "The number of scores must be equal to the number of labels"
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
