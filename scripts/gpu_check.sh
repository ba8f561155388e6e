#!/bin/bash
# whole GPU suite + smoke, nothing else (the cheapest "is the tree green" call)
mkdir -p gpurun_out
log=gpurun_out/check.log
: > $log
t0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu --timeout 600 >> $log 2>&1
echo "=== pytest exit $? ($(( $(date +%s) - t0 )) s)" >> $log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $log 2>&1
echo "=== smoke exit $?" >> $log
tail -15 $log
