#!/bin/bash
# One gpurun call: build check, smoke, GPU tests in isolated processes, bench lines, ncu launch list.
# Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh [quick|full]
mode=${1:-full}
mkdir -p gpurun_out
log=gpurun_out/round.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $log 2>&1
TMO=300 run python -c "import __graft_entry__ as g; g.smoke()"
TMO=300 run python -m pytest tests/test_gpu_tensor_scores.py -q -m gpu --timeout 120
TMO=600 run python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "kats or scan_path or bench_query or per_query or generated or remove_overwrite or duplicate or error_semantics or mask_pushdown"
TMO=600 run python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "tensor_path or device_search"
if [ "$mode" = "full" ]; then
  TMO=900 run python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 800 -k "full_size"
fi
for wl in c2 c1 c3a; do
  TMO=600 run python bench.py --workload $wl --steps 10 --warmup 3
done
tail -c 3000 $log
