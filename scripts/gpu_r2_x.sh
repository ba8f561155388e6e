#!/bin/bash
# rerank cut (select_kernel): whole GPU suite, C2 / C3b / C5 with the cut on and off, ncu launch list of the C2 bench command
mkdir -p gpurun_out
log=gpurun_out/r2x.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TMO=900 run python -m pytest tests -q -m gpu --timeout 600 -x
B="--no-cpu-baseline --secondary none --no-sustained"
for w in c2 c3b c5; do
  TMO=300 run python bench.py --workload $w --steps 20 --warmup 5 $B
  TMO=300 run python bench.py --workload $w --steps 20 --warmup 5 $B --opt rerank_cut=0
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2x_c2.csv \
  python bench.py --workload c2 --steps 2 --warmup 3 $B > gpurun_out/ncu_r2x.log 2>&1
echo "ncu exit $?" >> $log
grep -v "^{" $log | grep -v "^\[gemm" | tail -24
python scripts/benchsum.py $log | grep -v "^===\|clocks" | cut -c1-330
grep -E "select_kernel|rerank_finalize" gpurun_out/launches_r2x_c2.csv | head -6 | cut -c60-400
