#!/bin/bash
mkdir -p gpurun_out
python scripts/r2_probe.py cfg4 2> gpurun_out/t29.err | tail -1 | cut -c1-500
PROBE_BITS=6 PROBE_ROUNDS=8 python scripts/r2_probe.py tilesort 2>> gpurun_out/t29.err | grep -o '"RT_SORT_BITS.*"shade": [0-9.]*'
timeout 900 python -m pytest tests -x -q -m gpu -k "photon or knn or cfg4 or headline" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench29.json 2> gpurun_out/bench29.err; echo "bench rc=$?"; tail -2 gpurun_out/bench29.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench29.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"])
for k, v in d.get("extra", {}).items():
    print(k, v["value"], v["ms_per_step"], v.get("e2e", {}).get("value"), v["roofline"].get("frac"), v["roofline"].get("kernel"))
PY
