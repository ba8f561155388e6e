#!/bin/bash
# ncu captures (one call): launch list + full capture of the two hot kernels.
mkdir -p gpurun_out
log=gpurun_out/prof.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=300 run python -m pytest tests/test_gpu_tensor_scores.py -q -m gpu --timeout 120
TMO=300 run python scripts/prof_one.py --workload c2 --steps 3 --debug-sweep
TMO=300 run python scripts/prof_one.py --workload c5 --rows 4000000 --steps 3 --debug-sweep
for wl in c2 c3a; do TMO=600 run python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline; done
# ncu: plain run first (same command), then the capture
python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/plain_c2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 3 -c 1 \
    -o gpurun_out/prof_gemm_c2 python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 exit $?" >> $log
python scripts/prof_one.py --workload c3a --rows 4000000 --steps 2 > gpurun_out/plain_c3a.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 1 -c 1 \
    -o gpurun_out/prof_scan_c3a python scripts/prof_one.py --workload c3a --rows 4000000 --steps 2 > gpurun_out/ncu_c3a.log 2>&1
echo "ncu c3a exit $?" >> $log
tail -c 2500 $log
