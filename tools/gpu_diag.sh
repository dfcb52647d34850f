#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -X faulthandler -c "
import faulthandler, sys, runpy
faulthandler.dump_traceback_later(150, repeat=True)
sys.argv = ['bench.py', '--steps', '1', '--warmup', '1']
runpy.run_path('bench.py', run_name='__main__')
" > gpurun_out/bench_diag.log 2> gpurun_out/bench_diag.err; echo "bench exit $?"
tail -c 1500 gpurun_out/bench_diag.log; echo; grep -n "File\|line\|Thread" gpurun_out/bench_diag.err | head -40
