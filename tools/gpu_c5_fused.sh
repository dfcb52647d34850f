#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_similarity.py > gpurun_out/similarity.log 2>&1; echo "similarity exit $?"; tail -4 gpurun_out/similarity.log | cut -c1-400
export SS_RECSYS_VERBOSE=1 C5_REPS=3
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python tools/bench_c5.py 2000000 500000 0.05 > gpurun_out/c5_$name.log 2>&1; echo "$name exit $?"; grep "clusters" gpurun_out/c5_$name.log | tail -1; tail -1 gpurun_out/c5_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_all_reps'], d['spot_check_max_rel_err_top20'])"
  cp gpurun_out/c5.json gpurun_out/c5_$name.json
}
run 6x1024 SS_RECSYS_SHAPE=6x1024 SS_RECSYS_L2MB=128
run 6x1024_cl20 SS_RECSYS_SHAPE=6x1024 SS_RECSYS_L2MB=128 SS_RECSYS_CLUSTERS=20
run 7x1024 SS_RECSYS_SHAPE=7x1024 SS_RECSYS_L2MB=128
