#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=900 run python -m pytest tests -q -m gpu --timeout 600 -x
grep -E "^\{|exit [1-9]|passed|failed|rror|assert" $log | cut -c1-420
bash scripts/gpu_clients.sh > /dev/null 2>&1
cat gpurun_out/clients_probe.log | cut -c1-330
