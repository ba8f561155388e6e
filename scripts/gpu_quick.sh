#!/bin/bash
# quick correctness + timing loop: tests for both kernels, then a few timing points
mkdir -p gpurun_out
log=gpurun_out/quick.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=300 run python -m pytest tests/test_gpu_tensor_scores.py -q -m gpu --timeout 120
TMO=600 run python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "not full_size" -x
TMO=300 run python scripts/prof_one.py --workload c2 --steps 3 --debug-sweep
TMO=300 run python scripts/prof_one.py --workload c5 --rows 4000000 --steps 3
TMO=300 run python scripts/prof_one.py --workload c3a --rows 4000000 --steps 5
TMO=300 run python scripts/prof_one.py --workload c4 --rows 4000000 --steps 5
TMO=300 run python scripts/prof_one.py --workload c1 --steps 20
TMO=300 run python scripts/prof_one.py --workload c3b --rows 4000000 --steps 3
tail -c 3500 $log
