#!/bin/bash
# native concurrent-client probe (group commit on/off) for a few BASELINE shapes
mkdir -p gpurun_out
g++ -O2 -std=c++17 -pthread -I include scripts/probes/clients_probe.cpp -L vectordb-from-scratch_b200 -lgfi \
  -Wl,-rpath,$PWD/vectordb-from-scratch_b200 -o gpurun_out/clients_probe || exit 1
log=gpurun_out/clients_probe.log
: > $log
#                       metric rows     dim kind seed k  threads per_thread
timeout 300 gpurun_out/clients_probe 2 10000000 768 1 5 100 32 48 >> $log 2>&1   # C3a: dot, k=100
timeout 300 gpurun_out/clients_probe 0 10000000 384 0 6 10 32 48 >> $log 2>&1    # C4: L2
timeout 300 gpurun_out/clients_probe 1 1000000 768 1 3 10 64 64 >> $log 2>&1    # C2 rows: cosine
timeout 300 gpurun_out/clients_probe 0 10000 128 0 1 10 16 400 >> $log 2>&1     # C1
cat $log
