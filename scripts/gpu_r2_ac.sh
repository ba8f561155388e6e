#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c5.csv python scripts/prof_one.py --workload c5 --steps 1 > gpurun_out/ncu_launches_c5.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_c5.csv 2>/dev/null | tail -20 || tail -30 gpurun_out/launches_c5.csv
