#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
GFI_HOST_TRACE=1 timeout 200 python bench.py --steps 50 --warmup 3 --no-cpu-baseline >> $log 2>&1
grep "gfi trace" $log | tail -6
grep -E "^\{" $log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:40], 'step %.2f us  e2e %.2f us  kernel %.2f us' % (d['ms_per_step']*1e3, d['e2e']['ms_per_step']*1e3, d['roofline']['kernel_ms']*1e3))
"
