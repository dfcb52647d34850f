#!/bin/bash
# round-end style validation on one B200: full GPU suite, smoke, default bench line, reference arm, config timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_default.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'], d['check'], d['opt_in_f64_int8'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "reference exit $?"; tail -1 gpurun_out/bench_reference.log | cut -c1-300
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; echo "configs exit $?"; tail -3 gpurun_out/configs.log | cut -c1-400
