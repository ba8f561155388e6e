#!/bin/bash
# quick: C2 step + ncu launch list (select / rerank times)
mkdir -p gpurun_out
log=gpurun_out/r2z.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
B="--no-cpu-baseline --secondary none --no-sustained"
TMO=600 run python -m pytest tests/test_gpu_regressions_r2.py tests/test_gpu_certify.py tests/test_gpu_parity.py -q -k "not full_size" -m gpu --timeout 600 -x
TMO=300 run python bench.py --workload c2 --steps 20 --warmup 5 $B
TMO=300 run python bench.py --workload c5 --steps 20 --warmup 5 $B
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r2z_c2.csv \
  python bench.py --workload c2 --steps 2 --warmup 3 $B > gpurun_out/ncu_r2z.log 2>&1
echo "ncu exit $?" >> $log
grep -v "^{" $log | grep -v "^\[gemm" | tail -8
python scripts/benchsum.py $log | grep -v "^===\|clocks" | cut -c1-330
grep -E "select_kernel|rerank_finalize" gpurun_out/launches_r2z_c2.csv | head -4 | cut -c60-400
