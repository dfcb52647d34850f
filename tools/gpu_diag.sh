#!/bin/bash
mkdir -p gpurun_out
export SS_RECSYS_VERBOSE=1 C5_REPS=3
for u in 32 16; do
SS_RECSYS_UNIT=$u timeout 300 python tools/bench_c5.py 2000000 500000 0.05 > gpurun_out/c5_unit$u.log 2>&1; echo "unit $u exit $?"; grep clusters gpurun_out/c5_unit$u.log | tail -1; tail -1 gpurun_out/c5_unit$u.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_all_reps'])"
done
