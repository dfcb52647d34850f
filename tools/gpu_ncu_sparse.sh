#!/bin/bash
# ncu counters behind the dense <-> sparse switch (north star: "the choice justified by ncu counters"):
# both chains on the C3 shape at 5 % feature density (alpha = 0.95), where they tie.
mkdir -p gpurun_out
CMD="python tools/bench_sparse_dense.py 0.95"
$CMD > gpurun_out/sd_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:'spmm_kernel|ss_dgemm_kernel' -s 4 -c 4 \
    -o gpurun_out/r01_sparse_vs_dense_ncu -f $CMD > gpurun_out/ncu_sd.log 2>&1
echo "ncu exit $?"
