#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r2r.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }

TMO=600 run python bench.py --gpus 1 --steps 20 --warmup 5 --secondary c5 --no-sustained --no-cpu-baseline
grep -v "^{" $log | tail
python scripts/benchsum.py $log | grep -v "^===" | cut -c1-700
grep -o '"e2e": {[^}]*}[^}]*}' $log | cut -c1-600
