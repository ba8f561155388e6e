#!/bin/bash
# One gpurun call: smoke, the whole GPU suite, the default bench line, the reference arm and the filtered workloads.
mkdir -p gpurun_out
log=gpurun_out/verify.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv >> $log 2>&1
TMO=300 run python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 run python -m pytest tests -q -m gpu --timeout 800 -x
TMO=600 run python bench.py --steps 50 --warmup 3
for wl in "$@"; do TMO=600 run python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline; done
grep -E "passed|failed|exit|error" $log | tail -30
