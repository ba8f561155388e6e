#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=600 run python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "list_capacity or k1000 or per_query_k"
grep -E "^\{|exit [1-9]|passed|failed|rror|assert" $log | cut -c1-420
