#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -12 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_kernels.py > gpurun_out/kernels.log 2>&1; echo "kernels exit $?"; python - <<'PY'
import json
for k in json.load(open("gpurun_out/kernels.json"))["kernels"]:
    print(f'{k["kernel"][:60]:60s} {k["ms_median"]:9.3f} ms  {k["achieved_gbs"]:8.1f} GB/s  {k["frac_of_hbm_peak"]:.3f}')
PY
