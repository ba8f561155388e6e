#!/bin/bash
# round 2: bench at N=2 (in-process sharded handle under torchrun), both arms; A/B against the ranks variant
mkdir -p gpurun_out
log=gpurun_out/r2d.log
: > $log
N=${N:-2}
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
TMO=900 run $TR bench.py --gpus $N --steps 20 --warmup 5
TMO=600 run $TR bench.py --gpus $N --steps 200 --warmup 5 --secondary none --no-cpu-baseline
TMO=600 run $TR bench.py --gpus $N --steps 200 --warmup 5 --sharding ranks
grep -v "^{" $log | tail -40
