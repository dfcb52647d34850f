"""First product of the chain, T = (Xs' * (Y ./ ks)) ./ kf, in its two forms on C4's inner dimensions (Ns = Nf = 20 000,
Nt = 50 000): the dense DMMA GEMM (SS_T_FORM=dense) and the edge-list form of a sparse label matrix (csrc/ss_tsparse.cu,
SS_T_FORM=sparse), for a label density given by YD (default 0.05).  Per-launch times come from the library's own CUDA
events (ss_ctx_profile); the two results are compared entry by entry.  Writes gpurun_out/tform.json; run under gpurun."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simspread_b200 as ss
from simspread_b200._lib import SS_PREDICT_CLEAN, check

ctx = ss.Context(0)
L = ss.lib()
dev = torch.device("cuda:0")
dens = float(os.environ.get("YD", "0.05"))


def colmajor(rows, cols, fill):
    ld = (rows + 15) // 16 * 16
    buf = torch.empty((cols, ld), dtype=torch.float64, device=dev)
    for c0 in range(0, cols, 2000):
        fill(buf[c0:c0 + 2000])
    torch.cuda.synchronize()
    return buf, ss.DMat.wrap(ctx, buf.data_ptr(), rows, cols, ld)


uni = lambda b: b.copy_(torch.round(torch.rand(b.shape, device=dev, dtype=torch.float64) * 1e6) / 1e6)
bern = lambda b: b.copy_((torch.rand(b.shape, device=dev) < dens).to(torch.float64))
ns, nf, nt, nq = 20_000, 20_000, 50_000, 256  # few query rows: the second product is a small share of the call
bXs, mXs = colmajor(ns, nf, uni)
bY, mY = colmajor(ns, nt, bern)
bXq, mXq = colmajor(nq, nf, uni)
out, res = {"label_density": dens, "shape": {"ns": ns, "nf": nf, "nt": nt}}, {}
for form in ("dense", "sparse"):
    os.environ["SS_T_FORM"] = form
    bR, mR = colmajor(nq, nt, lambda b: b.zero_())
    for _ in range(2):  # the second call is the measured one
        ctx.profile(True)
        t0 = time.perf_counter()
        check(L.ss_predict_query(ctx.h, mXq.h, mXs.h, mY.h, mR.h, SS_PREDICT_CLEAN, None))
        wall = time.perf_counter() - t0
        prof = ctx.profile_read()
        ctx.profile(False)
    out[form] = {"call_wall_ms": wall * 1e3, "first_product_ms": prof[0][0], "first_product_flop": prof[0][1],
                 "second_product_ms": prof[1][0]}
    res[form] = bR[:, :nq].clone()
os.environ.pop("SS_T_FORM")
a, b = res["dense"], res["sparse"]
m = (a != -99) & (a != 0)
out["max_rel_diff_between_forms"] = float(((a - b).abs()[m] / a.abs()[m]).max())
out["clean_flags_equal"] = bool(torch.equal(a == -99, b == -99))
out["speedup_of_the_first_product"] = out["dense"]["first_product_ms"] / out["sparse"]["first_product_ms"]
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tform.json", "w"), indent=1)
print(json.dumps(out))
