#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_gpu.log
FULL="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$FULL > gpurun_out/bench_full1.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_full1.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['achieved'], d['roofline']['t_gemm_ms'], d['clocks'])"
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:dgemm_kernel -s 2 -c 2 --csv --log-file gpurun_out/r01_gemm_dram_full_v3.csv \
    $FULL > gpurun_out/ncu_dram.log 2>&1
echo "dram exit $?"; grep -E "dram__|duration|hit_rate" gpurun_out/r01_gemm_dram_full_v3.csv | awk -F'","' '{print $5, $13, $15}' | cut -c20-200
python - <<'PY'
import json,subprocess
# spread kernel timing only
PY
