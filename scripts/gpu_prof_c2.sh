#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_one.py --workload c2 --steps 3 --debug-sweep
python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/plain_c2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 3 -c 1 \
    -o gpurun_out/prof_gemm_c2 -f python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 exit $?"
