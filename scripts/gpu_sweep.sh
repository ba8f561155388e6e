#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
run python -m pytest tests/test_gpu_tensor_scores.py -q -m gpu --timeout 120
GFI_HOST_TRACE=1 run python scripts/probes/e2e_probe.py c2
bash scripts/gpu_launches.sh c2 >> $log 2>&1
grep -E "exit [1-9]|passed|failed|e2e ms|raw C" $log
grep "gfi trace" $log | tail -12
python scripts/launch_summary.py gpurun_out/launches_c2.csv 2>&1 | tail -30
