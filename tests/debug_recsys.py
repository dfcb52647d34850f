"""Debug aid for ss_recommend_topl: runs the (64 x 40000, L = 32) parity case and prints the rows whose
order differs from the oracle order, with the scores around the disagreement."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repository root
import simspread_b200 as ss
from oracle import simspread_oracle as o

users, items, dens, L = 64, 40000, 0.002, 32
rng = np.random.default_rng(users + items)
mask = rng.random((users, items)) < dens
for weighted in (True, False):
    Y = np.where(mask, np.round(rng.random((users, items)) + 0.5, 3), 0.0) if weighted else mask.astype(float)
    Y[3, :] = 0.0
    idx, val = ss.recommend_topl(Y, L)
    ks = np.count_nonzero(Y, axis=1)
    kt = np.count_nonzero(Y, axis=0)
    U = o._div_rows(np.ascontiguousarray(Y.T), kt) @ o._div_rows(Y, ks)
    F = Y @ U
    order = np.stack([o.sortperm_rev(F[u])[:L] for u in range(users)])
    bad = [u for u in range(users) if not np.array_equal(idx[u], order[u])]
    print("weighted", weighted, "rows differing:", bad)
    for u in bad[:4]:
        nz = np.count_nonzero(F[u])
        print(" row", u, "nonzero scores", nz)
        for r in range(L):
            if idx[u, r] != order[u, r]:
                print("  rank", r, "got", idx[u, r], F[u, idx[u, r]], val[u, r], "want", order[u, r], F[u, order[u, r]])
