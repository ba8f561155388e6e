#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=400 run python -m pytest tests/test_gpu_pairs.py -q -m gpu --timeout 120 -x
grep -E "^\{|exit [1-9]|passed|failed|rror|assert" $log | cut -c1-300
