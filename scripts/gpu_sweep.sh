#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/sweep.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-300} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=900 run python -m pytest tests -q -m gpu --timeout 600
TMO=300 run python -c "import __graft_entry__ as g; g.smoke()"
TMO=300 run python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5
TMO=300 run python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --workload c3a
grep -E "^\{|exit [1-9]|passed|failed|rror|assert|smoke ok" $log | cut -c1-330
