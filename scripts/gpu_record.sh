#!/bin/bash
# Round record: GPU tests, bench lines for every workload, ncu launch list of the default bench command,
# full ncu captures of the two hot kernels.  Everything lands in gpurun_out/ (summaries are copied to
# profiles/ by scripts/make_profiles.py).
mkdir -p gpurun_out
log=gpurun_out/record.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-900} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv >> $log 2>&1
TMO=900 run python -m pytest tests -q -m gpu --timeout 800
TMO=300 run python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 run python bench.py --steps 20 --warmup 3
TMO=900 run python bench.py --impl reference --steps 2 --warmup 1
for wl in c1 c3a c3b c4 c4f50 c4f1 c5; do TMO=900 run python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline; done
for wl in c3a c4; do TMO=900 run python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --opt tensor_auto=0; done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> $log
python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/plain_c2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 3 -c 1 \
    -o gpurun_out/prof_gemm_c2 -f python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/ncu_c2.log 2>&1
echo "ncu gemm c2 exit $?" >> $log
# C3a (q = 1 on 10M x 768): the cost model routes it to the tensor pass over the fp16 rows; the fp32 scan kernel is
# captured with that route switched off (it still serves filtered searches, small indexes and the fallback)
python scripts/prof_one.py --workload c3a --steps 2 > gpurun_out/plain_c3a.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 3 -c 1 \
    -o gpurun_out/prof_gemm_c3a -f python scripts/prof_one.py --workload c3a --steps 2 > gpurun_out/ncu_c3a.log 2>&1
echo "ncu gemm c3a exit $?" >> $log
python scripts/prof_one.py --workload c3a --steps 2 --opt tensor_auto=0 > gpurun_out/plain_c3a_scan.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 1 -c 1 \
    -o gpurun_out/prof_scan_c3a -f python scripts/prof_one.py --workload c3a --steps 2 --opt tensor_auto=0 > gpurun_out/ncu_c3a_scan.log 2>&1
echo "ncu scan c3a exit $?" >> $log
python scripts/prof_one.py --workload c5 --rows 2000000 --steps 2 > gpurun_out/plain_c5.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 3 -c 1 \
    -o gpurun_out/prof_gemm_c5 -f python scripts/prof_one.py --workload c5 --rows 2000000 --steps 2 > gpurun_out/ncu_c5.log 2>&1
echo "ncu gemm c5 exit $?" >> $log
python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/plain_rr.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rerank_finalize -s 2 -c 1 \
    -o gpurun_out/prof_rerank_c2 -f python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/ncu_rr.log 2>&1
echo "ncu rerank c2 exit $?" >> $log
# in-kernel cycle / clock diagnostics of the tensor pass (gemm_debug bit 5) with the epilogue, the loads and the
# MMAs switched off in turn (bits 2, 0-1, 4): the evidence for DESIGN.md section 5's power-wall / TMEM-port notes
python scripts/prof_one.py --workload c2 --steps 2 --debug-sweep > gpurun_out/clock_diag_c2.log 2>&1
python scripts/prof_one.py --workload c2 --steps 2 --debug-sweep --opt pair=1 > gpurun_out/clock_diag_c2_pair.log 2>&1
python scripts/prof_one.py --workload c5 --rows 4000000 --steps 2 --debug-sweep > gpurun_out/clock_diag_c5_4Mrows.log 2>&1
grep -E "passed|failed|exit" $log | tail -30
