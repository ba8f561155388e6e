#!/bin/bash
# round-2 call A: full GPU test-suite + K2 probe matrix (k-ring vs short-K kernel)
mkdir -p gpurun_out
log=gpurun_out/r2a.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $log 2>&1
TMO=600 run python -m pytest tests -m gpu -q --timeout 600 -x
TMO=400 run python scripts/probes/k2_probe.py 4000000 128 4096 euclidean
TMO=300 run python scripts/probes/k2_probe.py 4000000 128 4096 cosine 32,36
TMO=300 run python scripts/probes/k2_probe.py 12500000 128 4096 euclidean 32
TMO=300 run python scripts/probes/k2_probe.py 1000000 768 1024 cosine 32
tail -c 6000 $log
