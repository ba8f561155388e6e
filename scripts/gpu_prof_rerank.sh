#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/plain_rr.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rerank_finalize -s 2 -c 1 \
    -o gpurun_out/prof_rerank_c2 -f python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/ncu_rr.log 2>&1
echo "ncu rerank exit $?"
