#!/bin/bash
mkdir -p gpurun_out
C5_WEIGHTED=1 C5_REPS=3 SS_RECSYS_VERBOSE=1 timeout 400 python tools/bench_c5.py 2000000 500000 0.05 > gpurun_out/c5_weighted.log 2>&1; echo "weighted exit $?"; tail -1 gpurun_out/c5_weighted.log | cut -c1-800
cp gpurun_out/c5.json gpurun_out/c5_weighted.json
