#!/bin/bash
# ncu --set full of the C5 expand / extract kernels on the full-size graph (1 % of the users)
mkdir -p gpurun_out
CMD="python tools/bench_c5.py 2000000 500000 0.01"
$CMD > gpurun_out/c5_plain.log 2>&1 || { tail -5 gpurun_out/c5_plain.log; exit 1; }
tail -1 gpurun_out/c5_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rec_' -s 300 -c 6 \
    -o gpurun_out/r01_c5_ncu -f $CMD > gpurun_out/ncu_c5_full.log 2>&1
echo "ncu exit $?"
