#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
GFI_HOST_TRACE=1 TMO=200 run python bench.py --steps 50 --warmup 5
GFI_HOST_TRACE=1 TMO=120 run python scripts/prof_one.py --workload c2 --steps 5
grep -E "exit [1-9]|rror" $log | cut -c1-300
grep "gfi trace" $log | awk '{print}' | tail -22
grep -E "^\{" $log | cut -c1-120
