#!/bin/bash
# ncu evidence for round 1 (one GPU): launch list of the default bench command, one --set full
# capture of the dominant kernel (R = Xq*T) at 0.3 scale, DRAM traffic of the same kernel at full size.
mkdir -p gpurun_out
K='regex:dgemm_kernel|degrees_kernel|spread_kernel|featurize_kernel'
python bench.py > gpurun_out/bench_default.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv \
    --log-file gpurun_out/r01_launches.csv python bench.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"; tail -2 gpurun_out/bench_default.log | cut -c1-400
SMALL="python bench.py --scale 0.3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check"
$SMALL > gpurun_out/bench_s03.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_kernel -s 3 -c 1 \
    -o gpurun_out/r01_gemm_full -f $SMALL > gpurun_out/ncu_full.log 2>&1
echo "set full exit $?"
FULL="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-check"
$FULL > gpurun_out/bench_full1.log 2>&1 &&
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:dgemm_kernel -s 3 -c 1 --csv --log-file gpurun_out/r01_gemm_dram_full.csv \
    $FULL > gpurun_out/ncu_dram.log 2>&1
echo "dram exit $?"; tail -5 gpurun_out/r01_gemm_dram_full.csv
ls -la gpurun_out
