#!/bin/bash
mkdir -p gpurun_out
export SS_RECSYS_VERBOSE=1 C5_REPS=4
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python tools/bench_c5.py 2000000 500000 0.05 > gpurun_out/c5_$name.log 2>&1; echo "$name exit $?"; grep "clusters" gpurun_out/c5_$name.log | tail -1; tail -1 gpurun_out/c5_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_all_reps'], d['spot_check_max_rel_err_top20'])"
  cp gpurun_out/c5.json gpurun_out/c5_$name.json
}
run prefetch
run prefetch_unit8 SS_RECSYS_UNIT=8
timeout 300 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "recommender" > gpurun_out/pytest_rec.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_rec.log
