#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r2ab.log
: > $log
D=vectordb-from-scratch_b200
for rep in 1 2; do for v in old new; do
  cp $D/libgfi_$v.so $D/libgfi.so
  echo "=== $v" >> $log
  python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --secondary none --no-sustained >> $log 2>&1
done; done
cp $D/libgfi_new.so $D/libgfi.so
python scripts/benchsum.py $log | grep "===\|HEAD" | cut -c1-200
