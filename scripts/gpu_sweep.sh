#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=120 run python scripts/prof_one.py --workload c5 --rows 4000000 --steps 2 --opt gemm_debug=51
TMO=120 run python scripts/prof_one.py --workload c5 --rows 4000000 --steps 2 --opt gemm_debug=59
TMO=120 run python scripts/prof_one.py --workload c5 --rows 4000000 --steps 2 --opt gemm_debug=51 --metric cosine
TMO=120 run python scripts/prof_one.py --workload c5 --rows 4000000 --steps 2 --opt gemm_debug=59 --metric cosine
grep -E "^\{|exit [1-9]|passed|failed|rror|MHz" $log | grep -v "cycles [0-9]\{4,6\} " | cut -c1-300
