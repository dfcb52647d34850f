#!/bin/bash
# first GPU bring-up: probe, GEMM debug, parity tests, smoke, short + full bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 300 python tools/probe_box.py > gpurun_out/probe.log 2>&1; echo "probe exit $?"
timeout 120 python tools/debug_gemm.py > gpurun_out/debug_gemm.log 2>&1; echo "debug_gemm exit $?"
tail -30 gpurun_out/debug_gemm.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --scale 0.1 --steps 2 --warmup 1 > gpurun_out/bench_small.log 2>&1; echo "bench small exit $?"; tail -3 gpurun_out/bench_small.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_full.log 2>&1; echo "bench full exit $?"; tail -3 gpurun_out/bench_full.log
