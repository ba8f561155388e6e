#!/bin/bash
# A/B of the concurrent-client probe with library options taken from GFI_OPTS
mkdir -p gpurun_out
g++ -O2 -std=c++17 -pthread -I include scripts/probes/clients_probe.cpp -L vectordb-from-scratch_b200 -lgfi \
  -Wl,-rpath,$PWD/vectordb-from-scratch_b200 -o gpurun_out/clients_probe || exit 1
log=gpurun_out/clients_ab.log
: > $log
for opts in "" "zero_copy=0" "zero_copy=0,fused_tail=0"; do
  echo "## GFI_OPTS=$opts" >> $log
  GFI_OPTS=$opts timeout 300 gpurun_out/clients_probe 0 10000000 384 0 6 10 32 8 >> $log 2>&1
  GFI_OPTS=$opts timeout 300 gpurun_out/clients_probe 0 10000 128 0 1 10 16 400 >> $log 2>&1
done
python - <<PY
import json
for l in open("gpurun_out/clients_ab.log"):
    if l.startswith("#"): print(l.strip())
    if l.startswith("{"):
        j=json.loads(l); print(" ", j["rows"], j["dim"], "thr", j["threads"], "coalesce", j["coalesce"], "qps", j["qps"], "batches", j["coalesced_batches"], "err", j["errors"])
PY
