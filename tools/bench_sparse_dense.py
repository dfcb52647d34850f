"""Dense DMMA chain vs sparse row-split SpMM chain on the C3 shape (5000 queries, 5000 sources =
features, 2000 targets) across the alpha sweep: where does the switch belong?  Device time of
ss_predict_query vs ss_predict_query_csr (CSR build included separately).  Writes
gpurun_out/sparse_vs_dense.json; run under gpurun."""
import ctypes as C
import json
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simspread_b200 as ss
from simspread_b200._lib import SS_PREDICT_CLEAN, check

ctx = ss.Context(0)
ss.Context._default = ctx
L = ss.lib()
dev = torch.device("cuda:0")
ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)
nq, ns, nf, nt = 5000, 5000, 5000, 2000
rng = np.random.default_rng(20243)
Sq = np.round(rng.random((nq, nf)), 6)
Ssrc = np.round(rng.random((ns, nf)), 6)
Y = (rng.random((ns, nt)) < 0.01).astype(float)
dSq, dSs, dY = (ss.DMat.from_host(ctx, a) for a in (Sq, Ssrc, Y))
dXq, dXs = ss.DMat(ctx, nq, nf), ss.DMat(ctx, ns, nf)
R1, R2 = ss.DMat(ctx, nq, nt), ss.DMat(ctx, nq, nt)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        fn()
        e1.record(ext)
        ctx.sync()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


only = [float(a) for a in sys.argv[1:]] or [0.0, 0.5, 0.8, 0.9, 0.93, 0.95, 0.96, 0.97, 0.98, 0.99, 0.999]
out = []
for alpha in only:
    check(L.ss_featurize(ctx.h, dSq.h, alpha, 1, dXq.h))
    check(L.ss_featurize(ctx.h, dSs.h, alpha, 1, dXs.h))
    t_dense = timed(lambda: check(L.ss_predict_query(ctx.h, dXq.h, dXs.h, dY.h, R1.h, SS_PREDICT_CLEAN, None)))
    holder = {}

    def build():
        holder["cq"] = ss.DCsr.from_dense(ctx, dSq, alpha, True)                  # threshold fused with compaction
        holder["cs"] = ss.DCsr.from_dense(ctx, dSs, alpha, True, by_columns=True)

    t_build = timed(build, reps=3, warm=1)
    cq, cs = holder["cq"], holder["cs"]
    t_sparse = timed(lambda: check(L.ss_predict_query_csr(ctx.h, cq.h, cs.h, dY.h, R2.h, SS_PREDICT_CLEAN, None)))
    a, b = R1.to_host(), R2.to_host()
    nz = a != 0
    err = float(np.max(np.abs(a[nz] - b[nz]) / np.abs(a[nz]))) if nz.any() else 0.0
    pp = (cq.nnz + cs.nnz) * nt
    rec = {"alpha": alpha, "density_q": cq.density, "density_s": cs.density, "dense_ms": t_dense,
           "sparse_ms": t_sparse, "csr_build_ms": t_build, "partial_products": pp,
           "sparse_gbs_8B_per_pp": pp * 8 / (t_sparse * 1e-3) / 1e9,
           "dense_tflops": (2.0 * nf * nt * (ns + nq)) / (t_dense * 1e-3) / 1e12,
           "max_rel_diff_dense_vs_sparse": err, "faster": "sparse" if t_sparse < t_dense else "dense"}
    out.append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"shape": {"nq": nq, "ns": ns, "nf": nf, "nt": nt}, "points": out}, open("gpurun_out/sparse_vs_dense.json", "w"), indent=1)
