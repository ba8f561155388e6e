#!/bin/bash
# round 2: bench at N GPUs (in-process sharded handle under torchrun), both arms; sharded parity tests on all GPUs
mkdir -p gpurun_out
N=${N:-2}
log=gpurun_out/r2d_n$N.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
TMO=600 run python -m pytest tests/test_gpu_sharded.py -q -m gpu --timeout 300
TMO=900 run $TR bench.py --gpus $N --steps 20 --warmup 5
[ -z "$SKIP200" ] && TMO=600 run $TR bench.py --gpus $N --steps 200 --warmup 5 --secondary none --no-cpu-baseline --no-sustained
[ -n "$RANKS_AB" ] && TMO=600 run $TR bench.py --gpus $N --steps 200 --warmup 5 --sharding ranks
grep -v "^{" $log | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -40
