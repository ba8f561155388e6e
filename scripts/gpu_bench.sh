#!/bin/bash
# bench lines for the record + ncu launch list of the default bench command
mkdir -p gpurun_out
log=gpurun_out/bench.log
: > $log
run() { echo "=== $*" >> $log; timeout ${TMO:-900} "$@" >> $log 2>&1; echo "=== exit $?" >> $log; }
TMO=900 run python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 800 -k "full_size"
TMO=900 run python bench.py --steps 20 --warmup 3
TMO=900 run python bench.py --impl reference --steps 2 --warmup 1
for wl in c1 c3a c3b c4 c5; do TMO=900 run python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline; done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> $log
tail -c 1500 $log
