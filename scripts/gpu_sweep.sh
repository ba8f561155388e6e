#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=900 run python -m pytest tests -q -m gpu --timeout 600 -k "switches_off"
grep -E "^\{|exit [1-9]|passed|failed|rror|assert" $log | cut -c1-400
