#!/bin/bash
# fused cluster kernel for config 5: full GPU parity suite, bench at 5 % of the users, ncu of the fused kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
export SS_RECSYS_VERBOSE=1
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python tools/bench_c5.py 2000000 500000 0.05 > gpurun_out/c5_$name.log 2>&1; echo "$name exit $?"; grep "clusters" gpurun_out/c5_$name.log | tail -1; tail -1 gpurun_out/c5_$name.log | cut -c1-330
  cp gpurun_out/c5.json gpurun_out/c5_$name.json
}
run 8x1024 SS_RECSYS_SHAPE=8x1024
run 16x512 SS_RECSYS_SHAPE=16x512
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rec_fused' -s 1 -c 1 \
    -o gpurun_out/r01_c5_fused_ncu -f python tools/bench_c5.py 2000000 500000 0.002 > gpurun_out/ncu_c5_fused.log 2>&1
echo "ncu exit $?"
