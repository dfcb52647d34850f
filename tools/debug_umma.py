"""tcgen05 bring-up helper (run under gpurun with a timeout): integer-valued operands are exact in
TF32 / FP32, so any mismatch is a descriptor / layout / pipeline bug, not rounding."""
import sys

import numpy as np

sys.path.insert(0, ".")
import simspread_b200 as ss
from simspread_b200._lib import SS_OP_N, SS_OP_T, SS_PRECISION_TF32, check

ctx = ss.Context.default()


def run(A, B, op, prec):
    M, K = A.shape
    N = B.shape[1]
    dA = ss.DMat.from_host(ctx, A if op == SS_OP_N else A.T)
    dB, dC = ss.DMat.from_host(ctx, B), ss.DMat.from_host(ctx, np.full((M, N), -7.0))
    check(ss.lib().ss_gemm_lowp(ctx.h, op, dA.h, dB.h, dC.h, None, None, prec))
    return dC.to_host()


ok = True
for prec, pn in ((SS_PRECISION_TF32, "tf32"),):
    for op, nm in ((SS_OP_N, "N"), (SS_OP_T, "T")):
        for (M, N, K) in [(128, 256, 32), (128, 256, 128), (16, 8, 4), (300, 520, 260), (45, 664, 400), (129, 257, 33),
                          (1000, 1500, 2000)]:
            rng = np.random.default_rng(M + N + K)
            A = rng.integers(-3, 4, size=(M, K)).astype(float)
            B = rng.integers(-3, 4, size=(K, N)).astype(float)
            got, want = run(A, B, op, prec), A @ B
            bad = got != want
            if bad.any():
                ok = False
                r, c = np.nonzero(bad)
                print(f"[FAIL] {pn} op{nm} {M}x{N}x{K}: {bad.sum()}/{bad.size} wrong; rows {sorted(set(r))[:12]} "
                      f"cols {sorted(set(c))[:12]}; e.g. C[{r[0]},{c[0]}]={got[r[0], c[0]]} want {want[r[0], c[0]]}")
            else:
                print(f"[ok]   {pn} op{nm} {M}x{N}x{K}")
    # rounding behaviour on real data
    rng = np.random.default_rng(1)
    A, B = rng.random((500, 3000)), rng.random((3000, 700))
    got, want = run(A, B, SS_OP_N, prec), A @ B
    print(f"{pn}: max rel err on uniform data {np.max(np.abs(got - want) / want):.3e}")
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
