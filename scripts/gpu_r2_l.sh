#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r2l.log
: > $log
run() { echo "=== $*" >> $log; local t0=$(date +%s); timeout ${TMO:-600} "$@" >> $log 2>&1; echo "=== exit $? ($(( $(date +%s) - t0 )) s)" >> $log; }
TMO=300 run python scripts/probes/k2_probe.py 4000000 128 4096 cosine 32
TMO=300 run python scripts/probes/k2_probe.py 4000000 128 4096 euclidean 32
grep -v "^{" $log | grep -v "^\[gemm_topk\] cycles" | awk '/debug=32/{p=1} /^\[gemm_topk_sk\]/{c++} c<=24 || !/^\[gemm_topk_sk\]/' | tail -80
