#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_default.log | cut -c1-1500
timeout 900 python tools/bench_kernels.py > gpurun_out/kernels.log 2>&1; echo "kernels exit $?"; tail -8 gpurun_out/kernels.log | cut -c1-330
FULL="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check"
$FULL > gpurun_out/bench_full1.log 2>&1 &&
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:dgemm_kernel -s 2 -c 2 --csv --log-file gpurun_out/r01_gemm_dram_full_v2.csv \
    $FULL > gpurun_out/ncu_dram.log 2>&1
echo "dram exit $?"; grep -E "dram__|duration|hit_rate" gpurun_out/r01_gemm_dram_full_v2.csv | awk -F'","' '{print $5, $13, $15}' | cut -c1-200
