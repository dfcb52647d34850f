#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -8 gpurun_out/pytest_gpu.log
for P in tf32 f64_int8; do
  CMD="python bench.py --precision $P --scale 0.3 --steps 1 --warmup 1 --no-cpu-baseline --no-check"
  $CMD > gpurun_out/bench_${P}_s03.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ss_umma_kernel -s 3 -c 1 \
      -o gpurun_out/r01_umma_${P}_full -f $CMD > gpurun_out/ncu_umma_$P.log 2>&1
  echo "ncu $P exit $?"
done
ls -la gpurun_out/*.ncu-rep
