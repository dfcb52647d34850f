"""Where the wall time of a C2-sized cross_validate call goes (host-side cProfile, GPU box)."""
import cProfile
import os
import pstats
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simspread_b200 as ss

rng = np.random.default_rng(20241)
N, Nt = 445, 664
S = np.round(rng.beta(2, 5, size=(N, N)), 6)
np.fill_diagonal(S, 1.0)
Y = (rng.random((N, Nt)) < 0.0099).astype(float)
names = [f"D{i:04d}" for i in range(N)]
tn = [f"T{j:04d}" for j in range(Nt)]
DD, DT = ss.NamedArray(S, (names, names)), ss.NamedArray(Y, (names, tn))
for _ in range(3):
    ss.cross_validate(DT, DD, 0.35, weighted=False, k_=10, seed=1)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    ss.cross_validate(DT, DD, 0.35, weighted=False, k_=10, seed=1)
pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(28)
