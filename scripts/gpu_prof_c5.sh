#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_one.py --workload c5 --rows 2000000 --steps 2 > gpurun_out/plain_c5.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 3 -c 1 \
    -o gpurun_out/prof_gemm_c5 -f python scripts/prof_one.py --workload c5 --rows 2000000 --steps 2 > gpurun_out/ncu_c5.log 2>&1
echo "ncu c5 exit $?"
python scripts/prof_one.py --workload c5 --rows 2000000 --steps 3 --debug-sweep
