#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=400 run python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 -k "magnitude or error_semantics"
grep -E "^\{|exit [1-9]|passed|failed|rror|assert|Mismatch" $log | cut -c1-400 | head -30
