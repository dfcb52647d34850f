#!/bin/bash
# round-2 record on one B200: the bench line, its ncu launch list, per-kernel numbers, full parity suite
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json
timeout 600 python tools/bench_kernels.py > gpurun_out/kernels.log 2>&1; tail -3 gpurun_out/kernels.log
python bench.py --steps 1 --warmup 1 --no-side --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ss_|degrees|spread|featurize|gather|clean' -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-side --no-cpu-baseline --no-e2e > /dev/null 2>&1
wc -l gpurun_out/launches_bench.csv
