#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r2v.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TMO=900 run python -m pytest tests/test_gpu_regressions_r2.py tests/test_gpu_tensor_scores.py tests/test_gpu_parity.py -q -m gpu --timeout 600 -x -k "short_k or tensor or full_size_c5 or uniform"
TMO=600 run python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline --secondary none --no-sustained
TMO=600 run python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline --secondary none --no-sustained --opt short_k_seed=0
grep -v "^{" $log | grep -v "^\[gemm_topk" | tail -14
python scripts/benchsum.py $log | grep -v "^===" | cut -c1-330
