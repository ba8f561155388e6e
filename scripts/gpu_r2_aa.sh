#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r2aa.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TMO=900 run python -m pytest tests/test_gpu_regressions_r2.py tests/test_gpu_tensor_scores.py tests/test_gpu_parity.py -q -m gpu --timeout 600 -x -k "tensor or pair or full_size or route or mask"
TMO=600 run python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu-baseline --secondary c3b
grep -v "^{" $log | grep -v "^\[gemm_topk" | tail -8
python scripts/benchsum.py $log | grep -v "^===" | cut -c1-330
