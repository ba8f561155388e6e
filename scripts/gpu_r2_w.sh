#!/bin/bash
# rerank with unconditional loads: parity subset, C2 / C3b / C5 step times, ncu launch list of the C2 bench command
mkdir -p gpurun_out
log=gpurun_out/r2w.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TMO=600 run python -m pytest tests/test_gpu_parity.py tests/test_gpu_certify.py tests/test_gpu_regressions_r2.py -q -m gpu --timeout 600 -x -k "not full_size"
B="--no-cpu-baseline --secondary none --no-sustained"
TMO=300 run python bench.py --workload c2 --steps 20 --warmup 5 $B
TMO=300 run python bench.py --workload c3b --steps 20 --warmup 5 $B
TMO=300 run python bench.py --workload c5 --steps 10 --warmup 3 $B
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2w_c2.csv \
  python bench.py --workload c2 --steps 2 --warmup 3 $B > gpurun_out/ncu_r2w.log 2>&1
echo "ncu exit $?" >> $log
grep -v "^{" $log | tail -12
python scripts/benchsum.py $log | grep -v "^===" | cut -c1-330
grep -E "select_kernel|rerank_finalize|seed_finalize|prep_queries|convert_queries" gpurun_out/launches_r2w_c2.csv | head -12
