"""GEMM bring-up helper (run under gpurun): structured operands that expose fragment / swizzle /
permutation mistakes, with a compact description of where the result differs."""
import sys

import numpy as np

sys.path.insert(0, ".")
import simspread_b200 as ss
from simspread_b200._lib import SS_OP_N, SS_OP_T, check

ctx = ss.Context.default()


def run(A, B, op):
    M, K = A.shape
    N = B.shape[1]
    dA = ss.DMat.from_host(ctx, A if op == SS_OP_N else A.T)
    dB, dC = ss.DMat.from_host(ctx, B), ss.DMat(ctx, M, N)
    check(ss.lib().ss_gemm_f64(ctx.h, op, dA.h, dB.h, dC.h, None, None))
    return dC.to_host()


def describe(got, want, tag):
    bad = got != want
    if not bad.any():
        print(f"[ok]   {tag}")
        return True
    r, c = np.nonzero(bad)
    print(f"[FAIL] {tag}: {bad.sum()} / {bad.size} wrong; rows {sorted(set(r))[:20]} cols {sorted(set(c))[:20]}")
    for i in range(min(6, len(r))):
        print(f"        C[{r[i]},{c[i]}] = {got[r[i], c[i]]!r} want {want[r[i], c[i]]!r}")
    return False


ok = True
for op, nm in ((SS_OP_N, "N"), (SS_OP_T, "T")):
    for (M, N, K) in [(128, 128, 16), (128, 128, 32), (16, 8, 4), (256, 384, 160), (45, 664, 400), (129, 127, 17)]:
        rng = np.random.default_rng(M + N + K)
        # 1. A coded by (row, k), B = selector of one k: C[:, n] must be column k=n%K of A
        A = (np.arange(M)[:, None] * 100.0 + np.arange(K)[None, :])
        B = np.zeros((K, N))
        B[np.arange(N) % K, np.arange(N)] = 1.0
        ok &= describe(run(A, B, op), A @ B, f"op{nm} {M}x{N}x{K} coded-A x selector")
        # 2. random small integers
        A = rng.integers(-3, 4, size=(M, K)).astype(float)
        B = rng.integers(-3, 4, size=(K, N)).astype(float)
        ok &= describe(run(A, B, op), A @ B, f"op{nm} {M}x{N}x{K} random ints")
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
