"""Throughput of the similarity kernels fused with the threshold: ss_jaccard_featurize on real-valued
descriptors and ss_tanimoto_featurize_bits on bit-packed fingerprints, at C4-like pair counts.  Both are ALU-bound
(d operations per output element); reported as pair-descriptor operations per second and as output GB/s.
Writes gpurun_out/similarity.json; run under gpurun on one B200."""
import ctypes as C
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simspread_b200 as ss
from simspread_b200._lib import check

ctx = ss.Context(0)
L = ss.lib()
dev = torch.device("cuda:0")
ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)


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
    return min(ts), statistics.median(ts)


def colmajor(rows, cols):
    ld = (rows + 15) // 16 * 16
    buf = torch.rand((cols, ld), dtype=torch.float64, device=dev)
    return buf, ss.DMat.wrap(ctx, buf.data_ptr(), rows, cols, ld)


out = []
for na, nb, d in ((20000, 20000, 64), (50000, 20000, 1024)):
    bA, mA = colmajor(na, d)
    bB, mB = colmajor(nb, d)
    bX, mX = colmajor(na, nb)
    best, med = timed(lambda: check(L.ss_jaccard_featurize(ctx.h, mA.h, mB.h, 0.5, 1, mX.h)))
    ops = 6.0 * na * nb * d  # |a+b|, |a-b| (add + abs each folded), two accumulations: 6 FP64 operations per descriptor
    rec = {"kernel": "jaccard_featurize_kernel", "rows": na, "cols": nb, "descriptors": d, "ms_best": best, "ms_median": med,
           "pairs_per_s": na * nb / (med * 1e-3), "fp64_ops_per_s": ops / (med * 1e-3),
           "output_gbs": na * nb * 8 / (med * 1e-3) / 1e9}
    # spot check of one entry against the definition
    a, b = bA[:, 17], bB[:, 23]
    a1, a2 = ((a + b).abs() - (a - b).abs()).sum(), ((a + b).abs() + (a - b).abs()).sum()
    s = 1.0 - (1.0 - a1 / a2)
    want = s if s >= 0.5 else torch.zeros_like(s)
    rec["spot_check_abs_err"] = float((bX[23, 17] - want).abs().item())
    out.append(rec)
    print(json.dumps(rec), flush=True)
    del bA, bB, bX

for na, nb, words in ((100000, 20000, 32), (100000, 20000, 16)):  # 2048-bit and 1024-bit fingerprints
    FA = torch.randint(-2**62, 2**62, (na, words), dtype=torch.int64, device=dev) & torch.randint(-2**62, 2**62, (na, words), dtype=torch.int64, device=dev)
    FB = torch.randint(-2**62, 2**62, (nb, words), dtype=torch.int64, device=dev) & torch.randint(-2**62, 2**62, (nb, words), dtype=torch.int64, device=dev)
    bX, mX = colmajor(na, nb)
    best, med = timed(lambda: check(L.ss_tanimoto_featurize_bits(ctx.h, C.c_void_p(FA.data_ptr()), na, C.c_void_p(FB.data_ptr()), nb,
                                                                 words, 0.2, 1, mX.h)))
    rec = {"kernel": "tanimoto_bits_kernel", "rows": na, "cols": nb, "bits": 64 * words, "ms_best": best, "ms_median": med,
           "pairs_per_s": na * nb / (med * 1e-3), "word_ops_per_s": na * nb * words / (med * 1e-3),
           "output_gbs": na * nb * 8 / (med * 1e-3) / 1e9}
    out.append(rec)
    print(json.dumps(rec), flush=True)
    del FA, FB, bX

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/similarity.json", "w"), indent=1)
