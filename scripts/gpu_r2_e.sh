#!/bin/bash
# round 2: scan-path certification tests, then the whole GPU suite
mkdir -p gpurun_out
log=gpurun_out/r2e.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TMO=900 run python -m pytest tests/test_gpu_certify.py -q -m gpu --timeout 300
TMO=1500 run python -m pytest tests -q -m gpu --timeout 900 --deselect tests/test_gpu_certify.py
TMO=600 run python bench.py --gpus 1 --steps 20 --warmup 5 --no-sustained
grep -v "^{" $log | tail -60
