#!/bin/bash
# round 2: TMEM load-shape probe + host-side trace of the e2e call
mkdir -p gpurun_out
log=gpurun_out/r2i.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TMO=120 run scripts/probes/tmem_probe.bin 2000
export GFI_HOST_TRACE=1
TMO=300 run python scripts/probes/e2e_probe.py c2
cat $log | tail -120
