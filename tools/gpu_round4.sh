#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_sparse_dense.py > gpurun_out/sparse_vs_dense.log 2>&1; echo "sparse_dense exit $?"
cut -c1-420 gpurun_out/sparse_vs_dense.log | tail -12
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_full1.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_full1.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['achieved'], d['roofline']['t_gemm_ms'], d['e2e']['ms_per_step'], d['check'])"
