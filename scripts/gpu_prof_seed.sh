#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/plain_c2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 2 -c 1 \
    -o gpurun_out/prof_gemm_seed_c2 -f python scripts/prof_one.py --workload c2 --steps 2 > gpurun_out/ncu_seed.log 2>&1
echo "ncu seed exit $?"
python -m pytest tests -q -m gpu --timeout 900 -x -k "not full_size" 2>&1 | tail -2
python bench.py --steps 50 --warmup 3 --no-cpu-baseline | cut -c1-900
