#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_kernels.py > gpurun_out/kernels.log 2>&1; grep -i "csr" gpurun_out/kernels.log | cut -c1-400
