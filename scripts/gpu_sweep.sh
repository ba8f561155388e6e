#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=400 run python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor_scores.py -q -m gpu --timeout 120 -k "not full_size" -x
TMO=120 run python scripts/prof_one.py --workload c5 --rows 4000000 --steps 2 --debug-sweep
TMO=120 run python scripts/prof_one.py --workload c2 --steps 3 --debug-sweep
TMO=120 run python scripts/prof_one.py --workload c3b --rows 4000000 --steps 3
TMO=120 run python scripts/prof_one.py --workload c5 --rows 4000000 --steps 2 --metric cosine --opt gemm_debug=32
grep -E "^\{|exit [1-9]|passed|failed|rror|MHz" $log | grep -v "cycles [0-9]\{4,6\} " | cut -c1-300
