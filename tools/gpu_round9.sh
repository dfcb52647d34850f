#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "int8 or recommender or auc" > gpurun_out/pytest_gpu_sel.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu_sel.log
