#!/bin/bash
# The GPU-side command lines behind the numbers under profiles/ (run through `gpurun [--gpus N] -- 'bash tools/gpu_run.sh <what>'`).
#   tests    parity suite + smoke on one B200                      -> gpurun_out/pytest_gpu.log
#   bench    the bench line (N = 1), its ncu launch list            -> gpurun_out/bench_n1.json, launches_bench.csv
#   kernels  per-kernel numbers of the HBM-bound passes             -> gpurun_out/kernels.json
#   c5       sparse recommender: 5 % and 100 % of the users, ncu    -> gpurun_out/c5_*.json, c5_prof.ncu-rep
#   tform    first product: dense DMMA GEMM vs the edge-list form of a sparse label matrix -> gpurun_out/tform.json
#   prof     ncu --set full of the kernels changed late in round 2 (GEMM row bands on the C3 shape, CSR count / fill, top-L)
#   n2 / n8  multi-GPU: C-ABI sharded test (n2), bench line, C5 sharded by users, reference arm under torchrun (n8)
set -u
mkdir -p gpurun_out
case "${1:-tests}" in
tests)
  python __graft_entry__.py smoke 2>&1 | tail -2
  timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log ;;
bench)
  timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json
  python bench.py --steps 1 --warmup 1 --no-side --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ss_|degrees|spread|featurize|gather|clean' -c 200 --csv \
      --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-side --no-cpu-baseline --no-e2e > /dev/null 2>&1 ;;
kernels)
  timeout 600 python tools/bench_kernels.py > gpurun_out/kernels.log 2>&1; tail -20 gpurun_out/kernels.log | cut -c1-300 ;;
c5)
  C5_REPS=3 C5_OUT=gpurun_out/c5_stream_5pct.json timeout 300 python tools/bench_c5.py 2000000 500000 0.05 2>&1 | tail -1 | cut -c1-1200
  C5_REPS=2 C5_OUT=gpurun_out/c5_full.json timeout 300 python tools/bench_c5.py 2000000 500000 1.0 2>&1 | tail -1 | cut -c1-1200
  C5_OUT=gpurun_out/c5_small.json python tools/bench_c5.py 2000000 500000 0.01 > gpurun_out/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'tr_stream' -s 1 -c 1 -o gpurun_out/c5_prof \
      python tools/bench_c5.py 2000000 500000 0.01 > gpurun_out/ncu_full.log 2>&1 ;;
n2)
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sharded_c_abi" 2>&1 | tail -8 > gpurun_out/pytest_gpu_n2.log; cat gpurun_out/pytest_gpu_n2.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_n2.json ;;
n8)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -c 1500 gpurun_out/bench_n8.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 \
      tools/bench_multi.py --skip-c3 --skip-auc > gpurun_out/multi_n8.log 2>&1; tail -2 gpurun_out/multi_n8.log | cut -c1-600
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 \
      bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; tail -c 400 gpurun_out/bench_ref_n8.json ;;
tform)
  timeout 300 python tools/bench_tform.py 2>&1 | tail -1 | cut -c1-900 ;;
prof)
  python tools/profile_kernels_r02.py > gpurun_out/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'ss_dgemm_kernel|csr_count|csr_fill|topl_warp' -s 4 -c 4 -o gpurun_out/r02_late_kernels \
      python tools/profile_kernels_r02.py > gpurun_out/ncu_late.log 2>&1
  ncu -i gpurun_out/r02_late_kernels.ncu-rep --page details --csv > gpurun_out/r02_ncu_late_kernels_setfull_details.csv 2>/dev/null
  tail -2 gpurun_out/ncu_late.log ;;
*) echo "usage: tools/gpu_run.sh tests|bench|kernels|c5|n2|n8|prof|tform"; exit 2 ;;
esac
