#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench38.json 2> gpurun_out/bench38.err; echo "bench rc=$?"; tail -2 gpurun_out/bench38.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench38.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["rank0_ms"], d["parity"])
PY
