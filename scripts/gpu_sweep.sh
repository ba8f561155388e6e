#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=120 run python -m pytest tests/test_gpu_tensor_scores.py -q -m gpu --timeout 60
TMO=400 run python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 -k "not full_size" -x
TMO=120 run python scripts/prof_one.py --workload c2 --steps 5
TMO=120 run python scripts/prof_one.py --workload c2 --steps 5 --opt kp=32
TMO=120 run python scripts/prof_one.py --workload c3b --rows 4000000 --steps 3
TMO=120 run python scripts/prof_one.py --workload c5 --rows 4000000 --steps 3
TMO=200 run python bench.py --steps 50 --warmup 5
bash scripts/gpu_launches.sh c2 >> $log 2>&1
grep -E "^\{|exit [1-9]|passed|failed|rror|MHz" $log | cut -c1-600
python scripts/launch_summary.py gpurun_out/launches_c2.csv 2>&1 | tail -14
