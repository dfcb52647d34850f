#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}; shift
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/bench_multi.py "$@" > gpurun_out/multi_n$N.log 2>&1
echo "multi exit $?"; grep -v Warning gpurun_out/multi_n$N.log | tail -5 | cut -c1-700
