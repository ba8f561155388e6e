#!/bin/bash
# round 2: ncu --set full of the short-K main pass (C5 shape, 2M rows), L2 (coefficient epilogue) and cosine (raw)
mkdir -p gpurun_out
for m in euclidean cosine; do
python scripts/prof_one.py --workload c5 --rows 2000000 --steps 2 --metric $m > gpurun_out/plain_c5_$m.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_topk_sk -s 1 -c 1 \
    -o gpurun_out/r02_prof_sk_c5_$m -f python scripts/prof_one.py --workload c5 --rows 2000000 --steps 2 --metric $m > gpurun_out/ncu_c5_$m.log 2>&1
echo "ncu c5 $m exit $?"
done
