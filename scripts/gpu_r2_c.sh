#!/bin/bash
# round 2: the new bench line (headline + sustained + secondary) on one GPU
mkdir -p gpurun_out
log=gpurun_out/r2c.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TMO=900 run python bench.py --gpus 1 --steps 20 --warmup 5
tail -c 6000 $log
