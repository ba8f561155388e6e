#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r2t.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
for hits in 0 256 128; do
TMO=300 run python bench.py --workload c2 --steps 1000 --warmup 5 --no-cpu-baseline --secondary none --no-sustained --opt hits=$hits
done
TMO=300 run python bench.py --workload c2 --steps 1000 --warmup 5 --no-cpu-baseline --secondary none --no-sustained --opt hits=256 --opt kp=32
python scripts/benchsum.py $log | cut -c1-330
