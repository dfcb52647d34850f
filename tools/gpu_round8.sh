#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; echo "configs exit $?"; tail -3 gpurun_out/configs.log | cut -c1-500
timeout 600 python tools/bench_multi.py --skip-c5 --skip-auc > gpurun_out/multi_c3_n1.log 2>&1; echo "multi exit $?"; grep -v Warn gpurun_out/multi_c3_n1.log | tail -2 | cut -c1-700
cp gpurun_out/multi_n1.json gpurun_out/multi_c3_n1.json
