"""Per-kernel throughput of the HBM-bound passes of the path ("also report featurize
(elements/s, GB/s) and metric kernels (scores/s) individually").  CUDA events on the library
stream, best and median of 10 after 3 warm-ups; inputs are far larger than the 126 MB L2.
Writes gpurun_out/kernels.json; run under gpurun on one B200."""
import ctypes as C
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simspread_b200 as ss
from simspread_b200._lib import check

HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
ctx = ss.Context(0)
L = ss.lib()
dev = torch.device("cuda:0")
ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)


def colmajor(rows, cols, fill):
    ld = (rows + 15) // 16 * 16
    buf = torch.empty((cols, ld), dtype=torch.float64, device=dev)
    step = max(1, (1 << 27) // ld)
    for c0 in range(0, cols, step):
        fill(buf[c0:c0 + step])
    torch.cuda.synchronize()
    return buf, ss.DMat.wrap(ctx, buf.data_ptr(), rows, cols, ld)


def timed(fn, reps=10, warm=3):
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
    return min(ts), statistics.median(ts)


out = {"hbm_peak_gbs": HBM, "kernels": []}


def report(name, units, unit_name, bytes_alg, fn, note=""):
    best, med = timed(fn)
    rec = {"kernel": name, "units": units, "unit": unit_name, "ms_best": best, "ms_median": med,
           "units_per_s": units / (med * 1e-3), "algorithmic_bytes": bytes_alg,
           "achieved_gbs": bytes_alg / (med * 1e-3) / 1e9, "frac_of_hbm_peak": bytes_alg / (med * 1e-3) / 1e9 / HBM,
           "note": note}
    out["kernels"].append(rec)
    print(json.dumps(rec), flush=True)


uni = lambda b: b.copy_(torch.round(torch.rand(b.shape, device=dev, dtype=torch.float64) * 1e6) / 1e6)
bern = lambda b: b.copy_((torch.rand(b.shape, device=dev) < 0.05).to(torch.float64))

# featurize: C4-sized similarity block (100000 x 20000 = 16 GB), weighted, alpha 0.5, out of place
rows, cols = 100_000, 20_000
bS, mS = colmajor(rows, cols, uni)
bX, mX = colmajor(rows, cols, lambda b: b.zero_())
n = rows * cols
report("featurize_kernel (dense, weighted, alpha=0.5)", n, "elements", 16 * n,
       lambda: check(L.ss_featurize(ctx.h, mS.h, 0.5, 1, mX.h)), "8 B read + 8 B write per element")


def csr(alpha, weighted):
    h = C.c_void_p()
    check(L.ss_featurize_csr(ctx.h, mS.h, alpha, weighted, C.byref(h)))
    nnz = C.c_int64()
    check(L.ss_csr_info(h, None, None, C.byref(nnz), None))
    L.ss_csr_destroy(h)
    return nnz.value


nnz = csr(0.96, 1)
report("featurize_csr (count + keep-mask with lanes along rows, device-wide scan, mask-replay fill; weighted, alpha=0.96)", n, "elements",
       8 * n + 12 * nnz + 4 * rows,
       lambda: csr(0.96, 1), f"S read ONCE: 8 B per element + 12 B per kept edge ({nnz} edges, {nnz / n:.1%} dense); the fill pass "
       "replays the keep-mask of the count pass and touches only the sectors of kept entries; includes the host sync for nnz "
       "and cudaMallocAsync of the CSR arrays")
nnzb = csr(0.96, 0)
report("featurize_csr (same, binary graph: no values)", n, "elements", 8 * n + 4 * nnzb + 4 * rows, lambda: csr(0.96, 0),
       "the fill reads the mask only")
os.environ["SS_CSR_FILL"] = "tiled"
report("featurize_csr with the tiled fill forced (weighted, alpha=0.96)", n, "elements", 8 * n + 12 * nnz + 4 * rows,
       lambda: csr(0.96, 1), "same algorithmic bytes, S read twice")
del os.environ["SS_CSR_FILL"]
nnz9 = csr(0.8, 1)
report("featurize_csr (20 % dense: tiled fill; weighted, alpha=0.8)", n, "elements", 8 * n + 12 * nnz9 + 4 * rows,
       lambda: csr(0.8, 1), f"{nnz9} edges; above 10 % density the fill re-reads S through the transposing tiles")
del bX, mX

# degrees + spread on C4's Xs / Y
bXs, mXs = colmajor(20_000, 20_000, uni)
bY, mY = colmajor(20_000, 50_000, bern)
bW, mW = colmajor(20_000, 50_000, lambda b: b.zero_())
ks, kf, kt = ss.DIVec(ctx, 20_000), ss.DIVec(ctx, 20_000), ss.DIVec(ctx, 50_000)
n2 = 20_000 * 20_000 + 20_000 * 50_000
report("degrees_kernel x2 (ks, kf, kt of C4)", n2, "elements", 8 * n2 + 4 * 90_000,
       lambda: check(L.ss_degrees(ctx.h, mXs.h, mY.h, ks.h, kf.h, kt.h)), "8 B read per element")
n3 = 20_000 * 50_000
report("spread_kernel (Wst = Y ./ ks)", n3, "elements", 16 * n3,
       lambda: check(L.ss_spread_rows(ctx.h, mY.h, ks.h, mW.h)), "8 B read + 8 B write per element")
del bS, mS, bW, mW

# top-L and recall/precision@L on a C4-sized score matrix (100000 x 50000 = 40 GB)
bR, mR = colmajor(100_000, 50_000, uni)
idx = ss.DIVec(ctx, 20 * 100_000)
nr = 100_000 * 50_000
report("topl_kernel (L = 20)", nr, "scores", 8 * nr,
       lambda: check(L.ss_topl_rows(ctx.h, mR.h, 20, idx.h, None)), "8 B read per score")
del bR, mR

# AuROC / AuPRC on 10^8 (label, score) pairs (6-digit scores -> many ties)
M = 100_000_000
sc = torch.round(torch.rand(M, device=dev, dtype=torch.float64) * 1e6) / 1e6
lb = (torch.rand(M, device=dev) < 0.01).to(torch.uint8)
torch.cuda.synchronize()
res = (C.c_double * 2)()
report("auroc_auprc (radix sort + curve scan), M = 1e8", M, "scores", 9 * M * (2 + 3 * 8),
       lambda: check(L.ss_auroc_auprc(ctx.h, C.c_void_p(lb.data_ptr()), C.c_void_p(sc.data_ptr()), M, res)),
       "9 B per score x (key build + histogram + <= 8 passes x (upsweep read, downsweep read + write))")
out["auroc_auprc_value"] = [res[0], res[1]]
del sc, lb

# mid-size chain products (BASELINE config 3: 5000 x 2000 x 5000 -> 640 tiles = 4 waves + 48 tiles): the partial wave
# as whole tiles (SS_GEMM_TAIL_SPLIT=0) and as row bands (default); results are bit-identical (tests)
from simspread_b200._lib import SS_OP_N, SS_OP_T
out["gemm_mid_size"] = []
for (M_, N_, K_) in [(5000, 2000, 5000), (5000, 5000, 5000), (2000, 2000, 20000), (12500, 50000, 2500)]:
    bA, mA = colmajor(M_, K_, uni)
    bAt, mAt = colmajor(K_, M_, uni)
    bB, mB = colmajor(K_, N_, uni)
    bC, mC = colmajor(M_, N_, lambda b: b.zero_())
    for op, a in (("N", mA), ("T", mAt)):
        rec = {"M": M_, "N": N_, "K": K_, "opA": op, "tiles": ((M_ + 127) // 128) * ((N_ + 127) // 128)}
        for name, env in (("whole_tiles", "0"), ("row_bands", None)):
            if env is not None:
                os.environ["SS_GEMM_TAIL_SPLIT"] = env
            best, med = timed(lambda: check(L.ss_gemm_f64(ctx.h, SS_OP_N if op == "N" else SS_OP_T, a.h, mB.h, mC.h, None, None)))
            os.environ.pop("SS_GEMM_TAIL_SPLIT", None)
            rec[name] = {"ms_median": med, "ms_best": best, "tflops": 2.0 * M_ * N_ * K_ / (med * 1e-3) / 1e12}
        out["gemm_mid_size"].append(rec)
        print(json.dumps(rec), flush=True)
    del bA, mA, bAt, mAt, bB, mB, bC, mC

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/kernels.json", "w"), indent=1)
