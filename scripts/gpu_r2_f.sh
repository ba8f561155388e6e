#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r2f.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TMO=900 run python -m pytest tests/test_gpu_certify.py tests/test_gpu_c_client.py -q -m gpu --timeout 300
TMO=900 run python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -k "full_size or duplicate or above_the_kernels"
grep -v "^{" $log | tail -60
