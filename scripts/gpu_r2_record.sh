#!/bin/bash
# Round-2 record: GPU tests, smoke, the driver's bench command (both arms), ncu launch list of the bench command,
# full ncu captures of the hot kernels, in-kernel clock diagnostics.  Everything lands in gpurun_out/ (summaries are
# copied to profiles/ by scripts/make_profiles.py r02).
mkdir -p gpurun_out
log=gpurun_out/record.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-900} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv >> $log 2>&1
TMO=1200 run python -m pytest tests -q -m gpu --timeout 900
TMO=300 run python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 run python bench.py --gpus 1 --steps 20 --warmup 5
TMO=900 run python bench.py --impl reference --gpus 1 --steps 2 --warmup 1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --secondary none --no-sustained > gpurun_out/plain_bench.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --secondary none --no-sustained > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> $log
cap() {  # name, kernel regex, skip, prof_one args...
  local name=$1 rx=$2 skip=$3; shift 3
  python scripts/prof_one.py "$@" > gpurun_out/plain_$name.log 2>&1 && \
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
      -o gpurun_out/prof_$name -f python scripts/prof_one.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?" >> $log
}
cap gemm_c2 gemm_topk_kernel 3 --workload c2 --steps 2
cap gemm_c5 gemm_topk_sk 1 --workload c5 --steps 2
cap gemm_c3a gemm_topk_kernel 3 --workload c3a --steps 2
cap scan_c3a scan_topk 1 --workload c3a --steps 2 --opt tensor_auto=0
cap rerank_c2 rerank_finalize 2 --workload c2 --steps 2
python scripts/prof_one.py --workload c2 --steps 2 --debug-sweep > gpurun_out/clock_diag_c2.log 2>&1
python scripts/prof_one.py --workload c5 --steps 2 --opt gemm_debug=32 > gpurun_out/clock_diag_c5.log 2>&1
grep -E "passed|failed|exit" $log | tail -30
