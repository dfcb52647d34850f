#!/bin/bash
mkdir -p gpurun_out
SS_RECSYS_VERBOSE=1 SS_RECSYS_MODE=atomic C5_REPS=2 C5_OUT=gpurun_out/c5_atomic_5pct.json timeout 300 python tools/bench_c5.py 2000000 500000 0.05 2>&1 | grep -v Warn | tail -3 | cut -c1-600
SS_RECSYS_VERBOSE=1 SS_RECSYS_MODE=atomic C5_REPS=2 C5_DEGREES=pareto C5_OUT=gpurun_out/c5_atomic_pareto_1pct.json timeout 300 python tools/bench_c5.py 2000000 500000 0.01 2>&1 | grep -v Warn | tail -3 | cut -c1-600
