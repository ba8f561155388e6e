#!/bin/bash
# ncu launch lists (per-kernel device time) for a couple of workloads
mkdir -p gpurun_out
for wl in "$@"; do
  python scripts/prof_one.py --workload $wl --steps 2 > gpurun_out/plain_$wl.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_$wl.csv python scripts/prof_one.py --workload $wl --steps 2 > gpurun_out/ncu_l_$wl.log 2>&1
  echo "$wl exit $?"
done
