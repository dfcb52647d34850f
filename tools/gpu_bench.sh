#!/bin/bash
# full parity suite + the bench line (short run) on one B200
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 6000 gpurun_out/bench_n1.json
tail -5 gpurun_out/bench_n1.err
