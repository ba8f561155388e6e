#!/bin/bash
# last check of a round: smoke, the whole GPU suite, the default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -2
python bench.py > gpurun_out/final_default.log 2>&1
python scripts/benchlines.py gpurun_out/final_default.log
python - <<PY
import json
for l in open("gpurun_out/final_default.log"):
    if l.startswith("{"):
        j = json.loads(l); print(j["clocks"], j["steps"], j["cpu_baseline"])
PY
