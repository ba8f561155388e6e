#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=900 run python -m pytest tests -q -m gpu --timeout 600 -x
TMO=200 run python scripts/prof_one.py --workload c5 --steps 2
TMO=200 run python scripts/prof_one.py --workload c2 --steps 5
grep -E "^\{|exit [1-9]|passed|failed|rror|MHz" $log | cut -c1-300
