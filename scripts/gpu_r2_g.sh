#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r2g.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TMO=300 run python -m pytest tests/test_gpu_c_client.py -q -m gpu --timeout 300
TMO=1800 run python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 900 -k "full_size" --durations=8
grep -v "^{" $log | tail -60
