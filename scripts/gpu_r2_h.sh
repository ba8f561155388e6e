#!/bin/bash
# round 2: whole GPU suite + smoke + the driver's default bench line + ingest workload
mkdir -p gpurun_out
log=gpurun_out/r2h.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TMO=1500 run python -m pytest tests -q -m gpu --timeout 900 -x
TMO=300 run python -c "import __graft_entry__ as g; g.smoke()"
TMO=600 run python bench.py --workload ingest
TMO=900 run python bench.py --gpus 1 --steps 20 --warmup 5
grep -v "^{" $log | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -40
