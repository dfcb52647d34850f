#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "int8" > gpurun_out/pytest_int8.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_int8.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_default.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'], d['check'], d['opt_in_f64_int8'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
