#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "cross_validate or alpha_sweep" > gpurun_out/pytest_cv.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_cv.log
SS_SKIP_REFERENCE_FORMS=1 timeout 600 python tools/bench_configs.py > gpurun_out/configs_cv.log 2>&1; echo "configs exit $?"; head -1 gpurun_out/configs_cv.log | cut -c1-400
