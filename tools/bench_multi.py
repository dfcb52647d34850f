"""BASELINE configs 3 and 5 on N GPUs of one box (one process per GPU; launch with torchrun, or plain python for
N = 1).  Both shard without any data-path collective:
  C3: 21-point weighted alpha sweep on 5k queries x 5k sources x 2k targets, alpha points round-robin over ranks;
  C5: recommender-shaped sparse graph, 2M users x 500k items at 1e-4 density, fused top-20, user rows split
      evenly over the ranks (the graph, ~1.2 GB as two CSRs, is replicated).
Time = max over ranks of the device / wall time of the sharded call (all_reduce MAX).  Writes
gpurun_out/multi_n<N>.json on rank 0.   Usage: bench_multi.py [--c5-users 2000000] [--c5-frac 1.0] [--skip-c3]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--c5-users", type=int, default=2_000_000)
ap.add_argument("--c5-items", type=int, default=500_000)
ap.add_argument("--c5-frac", type=float, default=1.0)
ap.add_argument("--skip-c3", action="store_true")
ap.add_argument("--skip-c5", action="store_true")
ap.add_argument("--skip-auc", action="store_true")
args = ap.parse_args()

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

import simspread_b200 as ss  # noqa: E402
from simspread_b200._lib import check  # noqa: E402

ctx = ss.Context(local)
ss.Context._default = ctx
lib = ss.lib()


def max_over_ranks(x: float) -> float:
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


out = {"n_gpus": world}

# ---------------------------------------------------------------- C3
if not args.skip_c3:
    rng = np.random.default_rng(20243)
    nq, ns, nt = 5000, 5000, 2000
    Nn = nq + ns
    S3 = np.asfortranarray(np.round(rng.random((Nn, Nn)), 6))  # column-major, as a Julia Matrix{Float64} is
    Y3 = np.asfortranarray((rng.random((Nn, nt)) < 0.01).astype(float))
    n3 = [f"n{i}" for i in range(Nn)]
    t3 = [f"t{j}" for j in range(nt)]
    DD3, DT3 = ss.NamedArray(S3, (n3, n3)), ss.NamedArray(Y3, (n3, t3))
    alphas = [round(0.05 * i, 2) for i in range(21)]
    ss.alpha_sweep(DT3, DD3, n3[:nq], [0.0, 0.95])  # warm-up: both layouts
    res = {}
    for layout in ("auto", "dense"):
        barrier()
        t0 = time.perf_counter()
        tm = {}
        mine = ss.alpha_sweep(DT3, DD3, n3[:nq], alphas, rank=rank, world=world, layout=layout, timing=tm)
        torch.cuda.synchronize()
        t = max_over_ranks(time.perf_counter() - t0)
        t_setup, t_sweep = max_over_ranks(tm["setup_s"]), max_over_ranks(tm["sweep_s"])
        res[layout] = {"wall_s_max_over_ranks": t, "scores_per_s": 21 * nq * nt / t, "setup_s_max_over_ranks": t_setup,
                       "sweep_s_max_over_ranks": t_sweep, "sweep_scores_per_s": 21 * nq * nt / t_sweep}
        if layout == "auto":
            parts = [None] * world
            if world > 1:
                dist.all_gather_object(parts, mine)
            else:
                parts = [mine]
            pts = sorted((p for part in parts for p in part), key=lambda p: p["alpha"])
            res["points"] = pts
            res["layouts"] = {p["alpha"]: p["layout"] for p in pts}
    res["shape"] = {"nq": nq, "ns": ns, "nt": nt, "alphas": 21, "sharding": f"alpha points round-robin over {world} ranks"}
    res["includes"] = "per rank: upload of S and y, block gathers, per alpha featurize (dense or CSR), predict + clean!, AuROC/AuPRC, recall/precision@20, validity"
    out["C3_alpha_sweep"] = res
    if rank == 0:
        print(json.dumps({k: v for k, v in res.items() if k != "points"}), flush=True)
    del DD3, DT3, S3, Y3

# ---------------------------------------------------------------- C5
if not args.skip_c5:
    ns, nt, dens, L = args.c5_users, args.c5_items, 1e-4, 20
    g = torch.Generator(device=dev)
    g.manual_seed(20245)  # same graph on every rank
    deg = torch.poisson(torch.full((ns,), nt * dens, device=dev), generator=g).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(ns, device=dev), deg)
    cols = torch.randint(0, nt, (rows.numel(),), device=dev, generator=g)
    keys = torch.unique(rows * nt + cols)
    rows, cols = keys // nt, keys % nt
    nnz = keys.numel()
    y_ptr = torch.zeros(ns + 1, dtype=torch.int32, device=dev)
    y_ptr[1:] = torch.cumsum(torch.bincount(rows, minlength=ns), 0).to(torch.int32)
    y_idx = cols.to(torch.int32)
    keyt = torch.sort(cols * ns + rows).values
    yt_ptr = torch.zeros(nt + 1, dtype=torch.int32, device=dev)
    yt_ptr[1:] = torch.cumsum(torch.bincount(keyt // ns, minlength=nt), 0).to(torch.int32)
    yt_idx = (keyt % ns).to(torch.int32)
    del keys, keyt, rows, cols, deg
    torch.cuda.synchronize()

    def wrap(r, c, n, ptr, idx):
        h = C.c_void_p()
        check(lib.ss_csr_wrap(ctx.h, r, c, n, C.c_void_p(ptr.data_ptr()), C.c_void_p(idx.data_ptr()), None, C.byref(h)))
        return h

    hY, hYT = wrap(ns, nt, nnz, y_ptr, y_idx), wrap(nt, ns, nnz, yt_ptr, yt_idx)
    idx = torch.full((ns, L), -2, dtype=torch.int32, device=dev)
    val = torch.zeros((ns, L), dtype=torch.float64, device=dev)
    vi, vm = C.c_void_p(), C.c_void_p()
    check(lib.ss_ivec_wrap(ctx.h, C.c_void_p(idx.data_ptr()), ns * L, C.byref(vi)))
    check(lib.ss_mat_wrap(ctx.h, C.c_void_p(val.data_ptr()), L, ns, L, C.byref(vm)))
    n_proc = max(64 * world, int(ns * args.c5_frac))
    b, e = n_proc * rank // world, n_proc * (rank + 1) // world
    check(lib.ss_recommend_topl(ctx.h, hY, hYT, L, b, min(e, b + 2048), vi, vm))  # warm-up
    ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(ext)
    check(lib.ss_recommend_topl(ctx.h, hY, hYT, L, b, e, vi, vm))
    e1.record(ext)
    ctx.sync()
    ms = max_over_ranks(e0.elapsed_time(e1))
    # partial products of this rank's users: sum_{t' in Y[s]} sum_{s' in YT[t']} deg(s')
    ks = (y_ptr[1:] - y_ptr[:-1]).to(torch.float64)
    per_item = torch.zeros(nt, dtype=torch.float64, device=dev).index_add_(0, y_idx.long(), ks.repeat_interleave((y_ptr[1:] - y_ptr[:-1]).long()))
    pp = per_item[y_idx[int(y_ptr[b]): int(y_ptr[e])].long()].sum().reshape(1)
    if world > 1:
        dist.all_reduce(pp)
    pp = float(pp.item())
    # spot check on this rank: two users recomputed with torch sparse mat-vecs
    kt = (yt_ptr[1:] - yt_ptr[:-1]).to(torch.float64)
    Ysp = torch.sparse_csr_tensor(y_ptr.long(), y_idx.long(), torch.ones(nnz, dtype=torch.float64, device=dev), size=(ns, nt))
    YTsp = torch.sparse_csr_tensor(yt_ptr.long(), yt_idx.long(), torch.ones(nnz, dtype=torch.float64, device=dev), size=(nt, ns))
    worst = 0.0
    for s in (b, e - 1):
        a = torch.zeros(nt, dtype=torch.float64, device=dev)
        a[y_idx[y_ptr[s]:y_ptr[s + 1]].long()] = 1.0
        v2 = torch.mv(Ysp, torch.where(kt > 0, a / kt, torch.zeros_like(a)))
        F = torch.mv(YTsp, torch.where(ks > 0, v2 / ks, torch.zeros_like(v2)))
        want = torch.sort(F, descending=True).values[:L]
        worst = max(worst, float(((val[s] - want).abs() / want.abs().clamp_min(1e-300)).max().item()))
    worst = max_over_ranks(worst)
    out["C5_sparse_recommender"] = {
        "users": ns, "items": nt, "edges": nnz, "L": L, "users_processed": n_proc, "ms_max_over_ranks": ms,
        "scores_per_s": n_proc * nt / (ms * 1e-3), "partial_products": pp, "partial_products_per_s": pp / (ms * 1e-3),
        "sharding": f"user rows split over {world} ranks, graph replicated, no collective",
        "spot_check_max_rel_err_top20": worst}
    if rank == 0:
        print(json.dumps(out["C5_sparse_recommender"]), flush=True)

# ---------------------------------------------------------------- global AuROC / AuPRC over row-sharded scores
if not args.skip_auc:
    from simspread_b200.sharded import global_auroc_auprc
    m = 50_000_000 + 1000 * rank  # ragged slabs
    gen = torch.Generator(device=dev)
    gen.manual_seed(777 + rank)
    sc = torch.rand(m, dtype=torch.float64, device=dev, generator=gen)
    sc = torch.round(sc * 1e6) / 1e6  # ties across ranks
    lb = (torch.rand(m, device=dev, generator=gen) < 0.01 + 0.05 * sc).to(torch.float64)
    timings = {}
    for method in ("gather", "samplesort"):
        global_auroc_auprc(ss, ctx, torch, dist, lb, sc, world, rank, method=method)  # warm-up (buffers, NCCL channels)
        barrier()
        t0 = time.perf_counter()
        au = global_auroc_auprc(ss, ctx, torch, dist, lb, sc, world, rank, method=method)
        timings[method] = (max_over_ranks(time.perf_counter() - t0), au)
    t, au = timings["samplesort"]
    ok = None
    if rank == 0:  # the same slabs regenerated and concatenated on one GPU
        parts_s, parts_l = [], []
        for r_ in range(world):
            g2 = torch.Generator(device=dev)
            g2.manual_seed(777 + r_)
            m2 = 50_000_000 + 1000 * r_
            s2 = torch.round(torch.rand(m2, dtype=torch.float64, device=dev, generator=g2) * 1e6) / 1e6
            parts_l.append((torch.rand(m2, device=dev, generator=g2) < 0.01 + 0.05 * s2).to(torch.uint8))
            parts_s.append(s2)
        S_, L_ = torch.cat(parts_s[::-1]), torch.cat(parts_l[::-1])  # another order: the areas must not change
        ref = (C.c_double * 2)()
        torch.cuda.synchronize()
        check(lib.ss_auroc_auprc(ctx.h, C.c_void_p(L_.data_ptr()), C.c_void_p(S_.data_ptr()), int(S_.numel()), ref))
        ok = all(abs(ref[0] - a[0]) <= 1e-12 * abs(ref[0]) and abs(ref[1] - a[1]) <= 1e-12 * abs(ref[1])
                 for _, a in timings.values())
        out["global_auroc_auprc"] = {"scores_total": int(S_.numel()), "AuROC": au[0], "AuPRC": au[1],
                                     "samplesort_wall_s": t, "samplesort_scores_per_s": S_.numel() / t,
                                     "gather_wall_s": timings["gather"][0], "gather_scores_per_s": S_.numel() / timings["gather"][0],
                                     "matches_single_gpu_concatenation": bool(ok),
                                     "exchange": "samplesort: local radix sort, splitters from samples, one NCCL all-to-all of "
                                                 "(key, label) pairs, per-range integration, all-reduce of the partial areas; "
                                                 "gather: NCCL gather to rank 0 + single-GPU kernel"}
        print(json.dumps(out["global_auroc_auprc"]), flush=True)
        assert ok
    del sc, lb

if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/multi_n{world}.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
