"""BASELINE config 5 at full scale on one B200: recommender-shaped sparse graph, 2 000 000 users x
500 000 items, each entry Bernoulli(1e-4) (~1e8 edges), 2-layer NBI scores reduced on the fly to the
top-20 items per user (ss_recommend_topl).  The graph is generated on the device with torch and
wrapped (ss_csr_wrap); a few users are re-computed with torch sparse mat-vecs as a spot check.
Usage: python tools/bench_c5.py [users] [items] [user_fraction]   -> gpurun_out/c5.json"""
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import simspread_b200 as ss
from simspread_b200._lib import check

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
frac = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
dens, L = 1e-4, 20
dev = torch.device("cuda:0")
ctx = ss.Context(0)
lib = ss.lib()
from c5_graph import make_graph, partial_products, wrap  # noqa: E402

WEIGHTED = os.environ.get("C5_WEIGHTED") == "1"  # ratings in [0.5, 1.5) instead of 0/1 edges
G = make_graph(ns, nt, dens, dev, os.environ.get("C5_DEGREES", "poisson"), WEIGHTED)
nnz, y_ptr, y_idx, y_val, yt_ptr, yt_idx, yt_val = (G[k_] for k_ in ("nnz", "y_ptr", "y_idx", "y_val", "yt_ptr", "yt_idx", "yt_val"))
hY, hYT = wrap(lib, ctx, check, G)
idx = torch.full((ns, L), -2, dtype=torch.int32, device=dev)
val = torch.zeros((ns, L), dtype=torch.float64, device=dev)
vi, vm = C.c_void_p(), C.c_void_p()
check(lib.ss_ivec_wrap(ctx.h, C.c_void_p(idx.data_ptr()), ns * L, C.byref(vi)))
check(lib.ss_mat_wrap(ctx.h, C.c_void_p(val.data_ptr()), L, ns, L, C.byref(vm)))
s_end = max(64, int(ns * frac))
ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = int(os.environ.get("C5_REPS", "1"))
all_ms = []
build = None
ATOMIC = os.environ.get("SS_RECSYS_MODE") in ("atomic", "groups")
if ATOMIC:   # the round-1 kernels: one call does everything
    check(lib.ss_recommend_topl(ctx.h, hY, hYT, L, 0, min(2048, s_end), vi, vm))  # warm-up
    for _ in range(reps):
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        e0.record(ext)
        check(lib.ss_recommend_topl(ctx.h, hY, hYT, L, 0, s_end, vi, vm))
        e1.record(ext)
        ctx.sync()
        wall = time.perf_counter() - t0
        all_ms.append(e0.elapsed_time(e1))
if not ATOMIC:   # does the transfer matrix fit?  (heavy-tailed graphs: no -> one call, the library picks the kernel)
    hU = C.c_void_p()
    t0 = time.perf_counter()
    e0.record(ext)
    if lib.ss_transfer_build(ctx.h, hY, hYT, C.byref(hU)) != 0:
        ATOMIC = "declined"
if ATOMIC == "declined":
    check(lib.ss_recommend_topl(ctx.h, hY, hYT, L, 0, min(2048, s_end), vi, vm))  # warm-up
    for _ in range(reps):
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        e0.record(ext)
        check(lib.ss_recommend_topl(ctx.h, hY, hYT, L, 0, s_end, vi, vm))
        e1.record(ext)
        ctx.sync()
        wall = time.perf_counter() - t0
        all_ms.append(e0.elapsed_time(e1))
elif not ATOMIC:        # transfer matrix U built once (timed on its own), then streamed per source range
    e1.record(ext)
    ctx.sync()
    info = (C.c_int64 * 4)()
    check(lib.ss_transfer_info(hU, info))
    build = {"ms": e0.elapsed_time(e1), "wall_s": time.perf_counter() - t0, "entries": int(info[0]), "bytes": int(info[1]),
             "tile_width": int(info[2]), "tiles": int(info[3])}
    check(lib.ss_recommend_topl_transfer(ctx.h, hY, hU, L, 0, min(2048, s_end), vi, vm))  # warm-up
    for _ in range(reps):
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        e0.record(ext)
        check(lib.ss_recommend_topl_transfer(ctx.h, hY, hU, L, 0, s_end, vi, vm))
        e1.record(ext)
        ctx.sync()
        wall = time.perf_counter() - t0
        all_ms.append(e0.elapsed_time(e1))
    idx_first = idx[:s_end].clone()
    val_first = val[:s_end].clone()
    check(lib.ss_recommend_topl_transfer(ctx.h, hY, hU, L, 0, s_end, vi, vm))
    ctx.sync()
    build["second_run_bit_identical"] = bool(torch.equal(idx_first, idx[:s_end]) and
                                             torch.equal(val_first.view(torch.int64), val[:s_end].view(torch.int64)))
    check(lib.ss_transfer_destroy(hU))
ms = sorted(all_ms)[len(all_ms) // 2]
pp = partial_products(G, s_end)
ks = (y_ptr[1:] - y_ptr[:-1]).to(torch.float64)
kt = (yt_ptr[1:] - yt_ptr[:-1]).to(torch.float64)
# spot check: three users recomputed with torch sparse mat-vecs
ones = torch.ones(nnz, dtype=torch.float64, device=dev)
Ysp = torch.sparse_csr_tensor(y_ptr.long(), y_idx.long(), y_val if WEIGHTED else ones, size=(ns, nt))
YTsp = torch.sparse_csr_tensor(yt_ptr.long(), yt_idx.long(), yt_val if WEIGHTED else ones, size=(nt, ns))
worst = 0.0
for s in (0, s_end // 2, s_end - 1):
    a = torch.zeros(nt, dtype=torch.float64, device=dev)
    a[y_idx[y_ptr[s]:y_ptr[s + 1]].long()] = y_val[y_ptr[s]:y_ptr[s + 1]] if WEIGHTED else 1.0
    v1 = torch.where(kt > 0, a / kt, torch.zeros_like(a))
    v2 = torch.mv(Ysp, v1)
    F = torch.mv(YTsp, torch.where(ks > 0, v2 / ks, torch.zeros_like(v2)))
    want = torch.sort(F, descending=True).values[:L]
    got = val[s]
    worst = max(worst, float(((got - want).abs() / want.abs().clamp_min(1e-300)).max().item()))
    assert bool((F[idx[s].long()] - got).abs().max() <= 1e-12 * got.abs().max().clamp_min(1e-300))
out = {"users": ns, "items": nt, "edges": nnz, "L": L, "degrees": os.environ.get("C5_DEGREES", "poisson"), "weighted": WEIGHTED,
       "max_user_degree": int((y_ptr[1:] - y_ptr[:-1]).max().item()), "max_item_degree": int((yt_ptr[1:] - yt_ptr[:-1]).max().item()), "users_processed": s_end, "ms": ms, "wall_s": wall,
       "scores_per_s": s_end * nt / (ms * 1e-3), "partial_products": pp,
       "partial_products_per_s": pp / (ms * 1e-3), "achieved_gbs_4B_per_pp": pp * 4 / (ms * 1e-3) / 1e9,
       "achieved_gbs_10B_per_pp": pp * 10 / (ms * 1e-3) / 1e9, "mode": os.environ.get("SS_RECSYS_MODE", "stream") if ATOMIC != "declined" else "stream declined (transfer matrix too large) -> atomic",
       "tile": os.environ.get("SS_RECSYS_TILE", "2048"), "transfer_build": build,
       "kernel_launches": ctx.launch_count() - l0, "ms_all_reps": all_ms, "spot_check_max_rel_err_top20": worst}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(os.environ.get("C5_OUT", "gpurun_out/c5.json"), "w"), indent=1)
