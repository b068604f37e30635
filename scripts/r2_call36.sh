#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench36.json 2> gpurun_out/bench36.err; echo "bench rc=$?"; tail -2 gpurun_out/bench36.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench36.json") if l.startswith("{")][-1])
print(d["value"], d["e2e"]["value"])
print(json.dumps(d["extra"]["cfg4_end_to_end"], indent=1))
PY
