#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=900 run python -m pytest tests -q -m gpu --timeout 600 -k "full_size or mask or filter"
TMO=200 run python bench.py --workload c4f50 --steps 20 --warmup 3 --no-cpu-baseline
TMO=200 run python bench.py --workload c4f1 --steps 20 --warmup 3 --no-cpu-baseline
grep -E "^\{|exit [1-9]|passed|failed|rror|assert" $log | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['config']['workload'][:40], 'step %.2f us  e2e %.2f us  kernel %.2f us' % (d['ms_per_step']*1e3, d['e2e']['ms_per_step']*1e3, d['roofline']['kernel_ms']*1e3), d['roofline']['kernel'][:12], d['roofline']['bound'], '%.3f' % d['roofline']['frac'], 'fb', d['fallback_queries'])
    else: print(l.strip()[:300])
"
