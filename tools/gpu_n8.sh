#!/bin/bash
# eight GPUs: the bench line at N = 8 (C-ABI communicator, sharded Xs upload), the reference arm under torchrun, C5 sharded by users
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -c 2500 gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/bench_multi.py --skip-c3 --skip-auc > gpurun_out/multi_n8.log 2>&1; tail -3 gpurun_out/multi_n8.log | cut -c1-900
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; tail -c 700 gpurun_out/bench_ref_n8.json
