#!/bin/bash
# two-GPU sanity check of the closing tree: sharded parity tests (one shard per GPU) + the driver's N=2 headline
mkdir -p gpurun_out
log=gpurun_out/n2.log
: > $log
timeout 100 python -m pytest tests/test_gpu_sharded.py -q -m gpu --timeout 90 -x >> $log 2>&1
echo "=== pytest exit $?" >> $log
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --secondary none --no-sustained >> $log 2>&1
echo "=== bench exit $?" >> $log
grep -E "passed|failed|exit" $log
python scripts/benchsum.py $log | grep -v "^===" | cut -c1-330
