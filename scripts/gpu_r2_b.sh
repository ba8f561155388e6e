#!/bin/bash
# round 2: sharded handle (in-process, several GPUs) -- correctness, then the whole GPU suite after the api.cu split
mkdir -p gpurun_out
log=gpurun_out/r2b.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
nvidia-smi -L >> $log
TMO=900 run python -m pytest tests/test_gpu_sharded.py -q -m gpu --timeout 300
TMO=900 run python -m pytest tests -q -m gpu --timeout 600 -k "not sharded"
tail -c 3000 $log
