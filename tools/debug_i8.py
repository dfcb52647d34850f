"""Bring-up of the INT8-sliced FP64 GEMM (run under gpurun with a timeout)."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import simspread_b200 as ss
from simspread_b200._lib import SS_OP_N, SS_OP_T, SS_PRECISION_F64_INT8, check

ctx = ss.Context.default()


def run(A, B, op):
    M, K = A.shape
    N = B.shape[1]
    dA = ss.DMat.from_host(ctx, A if op == SS_OP_N else A.T)
    dB, dC = ss.DMat.from_host(ctx, B), ss.DMat.from_host(ctx, np.full((M, N), -7.0))
    check(ss.lib().ss_gemm_lowp(ctx.h, op, dA.h, dB.h, dC.h, None, None, SS_PRECISION_F64_INT8))
    return dC.to_host()


ok = True
for op, nm in ((SS_OP_N, "N"), (SS_OP_T, "T")):
    for (M, N, K) in [(128, 256, 128), (16, 8, 4), (300, 520, 260), (129, 257, 133), (1000, 1500, 2000), (200, 300, 20000)]:
        rng = np.random.default_rng(M + N + K)
        # 1. small integers: every slice product is exact and the answer is an integer
        A = rng.integers(0, 200, size=(M, K)).astype(float)
        B = rng.integers(0, 200, size=(K, N)).astype(float)
        got, want = run(A, B, op), A @ B
        bad = got != want
        tag = f"op{nm} {M}x{N}x{K}"
        if bad.any():
            ok = False
            r, c = np.nonzero(bad)
            print(f"[FAIL] ints {tag}: {bad.sum()}/{bad.size} wrong; rows {sorted(set(r))[:10]} cols {sorted(set(c))[:10]}; "
                  f"C[{r[0]},{c[0]}]={got[r[0], c[0]]!r} want {want[r[0], c[0]]!r}")
        else:
            print(f"[ok]   ints {tag}")
        # 2. saturating digits: all-255 slices stress the unsigned 32-bit accumulation
        A = np.full((M, K), 1.0 - 2.0 ** -48)
        B = np.full((K, N), 1.0 - 2.0 ** -48)
        got = run(A, B, op)
        want = np.full((M, N), float(K) * (1.0 - 2.0 ** -48) ** 2)
        e = np.max(np.abs(got - want) / want)
        print(f"       all-ones digits {tag}: max rel err {e:.3e}")
        ok &= e < 1e-12
        # 3. uniform data
        A, B = rng.random((M, K)), rng.random((K, N))
        got = run(A, B, op)
        want = (A.astype(np.longdouble) @ B.astype(np.longdouble)).astype(float)
        e = np.max(np.abs(got - want) / want)
        print(f"       uniform {tag}: max rel err {e:.3e}")
        ok &= e < 1e-12
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
