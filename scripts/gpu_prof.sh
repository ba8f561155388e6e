#!/bin/bash
# ncu captures (one call): full capture of the two hot kernels + timing points.
mkdir -p gpurun_out
log=gpurun_out/prof.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=600 run python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "tensor_path or scan_path" -x
TMO=300 run python scripts/prof_one.py --workload c2 --steps 3 --debug-sweep
TMO=300 run python scripts/prof_one.py --workload c5 --rows 4000000 --steps 3
TMO=300 run python scripts/prof_one.py --workload c3b --rows 4000000 --steps 3
TMO=300 run python scripts/prof_one.py --workload c3a --rows 4000000 --steps 5
TMO=300 run python scripts/prof_one.py --workload c3a --rows 4000000 --steps 5 --opt scan_stages=2
TMO=300 run python scripts/prof_one.py --workload c3a --rows 4000000 --steps 5 --opt scan_stages=3
TMO=300 run python scripts/prof_one.py --workload c4 --rows 4000000 --steps 5
python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/plain_c2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 3 -c 1 \
    -o gpurun_out/prof_gemm_c2 -f python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 exit $?" >> $log
python scripts/prof_one.py --workload c3a --rows 4000000 --steps 2 > gpurun_out/plain_c3a.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 1 -c 1 \
    -o gpurun_out/prof_scan_c3a -f python scripts/prof_one.py --workload c3a --rows 4000000 --steps 2 > gpurun_out/ncu_c3a.log 2>&1
echo "ncu c3a exit $?" >> $log
tail -c 3000 $log
