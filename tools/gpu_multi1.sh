#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "alpha_sweep" > gpurun_out/pytest_sweep.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_sweep.log
timeout 900 python tools/bench_multi.py > gpurun_out/multi_n1.log 2>&1; echo "multi exit $?"; tail -4 gpurun_out/multi_n1.log | cut -c1-600
