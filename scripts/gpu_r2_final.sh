#!/bin/bash
# Round-2 closing record: whole GPU suite, smoke, the driver's bench command, ncu launch list of the bench command.
mkdir -p gpurun_out
log=gpurun_out/final.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-900} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv >> $log 2>&1
TMO=900 run python -m pytest tests -q -m gpu --timeout 600
TMO=120 run python -c "import __graft_entry__ as g; g.smoke()"
TMO=600 run python bench.py --gpus 1 --steps 20 --warmup 5
B="--no-cpu-baseline --secondary none --no-sustained"
python bench.py --steps 2 --warmup 3 $B > gpurun_out/plain_bench.log 2>&1 && \
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_bench_c2.csv python bench.py --steps 2 --warmup 3 $B > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> $log
grep -E "passed|failed|exit|smoke" $log | cut -c1-200 | tail -12
python scripts/benchsum.py $log | grep -v "^===" | cut -c1-330
