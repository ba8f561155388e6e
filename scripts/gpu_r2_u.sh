#!/bin/bash
mkdir -p gpurun_out
python scripts/probes/fallback_counts.py > gpurun_out/r2u_fallbacks.log 2>&1; tail -8 gpurun_out/r2u_fallbacks.log
python scripts/prof_one.py --workload c5 --steps 2 --opt gemm_debug=32 > gpurun_out/r2u_diag.log 2>&1
echo "diag exit $?"
python scripts/prof_one.py --workload c5 --steps 2 > gpurun_out/plain_c5_full.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_topk_sk -s 1 -c 1 \
    -o gpurun_out/r02_prof_sk_c5_full -f python scripts/prof_one.py --workload c5 --steps 2 > gpurun_out/ncu_c5_full.log 2>&1
echo "ncu exit $?"
grep "gemm_topk_sk" gpurun_out/r2u_diag.log | tail -12
