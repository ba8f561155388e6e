#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=900 run python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -x -k "not full_size"
bash scripts/gpu_launches.sh c2 c3a c1 >> $log 2>&1
grep -E "^\{|exit [1-9]|passed|failed|rror|MHz" $log | cut -c1-300
for w in c2 c3a c1; do python scripts/launch_summary.py gpurun_out/launches_$w.csv 2>&1 | grep -E "select|rerank" | grep -v last; done
