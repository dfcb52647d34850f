#!/bin/bash
# two GPUs: the C-ABI sharded test (two processes, file rendezvous) and the bench line at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sharded_c_abi" 2>&1 | tail -8 > gpurun_out/pytest_gpu_n2.log
cat gpurun_out/pytest_gpu_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 3500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
