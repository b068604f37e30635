#!/bin/bash
# what the driver runs at round end, on the final commit: GPU tests, smoke, default bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench52.json 2> gpurun_out/bench52.err; echo "bench rc=$?"; tail -2 gpurun_out/bench52.err
timeout 900 python bench.py --impl reference > gpurun_out/bench52_ref.json 2>> gpurun_out/bench52.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench52.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "gpu_launches", "dtype", "scaling")})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "roofline", d["roofline"]["frac"], d["roofline"]["bound"], "clocks", d["clocks"])
print("cpu_baseline", {k: d["cpu_baseline"][k] for k in ("value", "cores", "kind")})
for k, v in d.get("extra", {}).items():
    if "value" in v: print(k, v["value"], v["ms_per_step"], v["e2e"]["value"], v["roofline"]["frac"], v["parity"].get("window_identical_frac"))
    else: print(k, v.get("ms_per_call"), v.get("speedup_vs_reference_program"))
r = json.loads([l for l in open("gpurun_out/bench52_ref.json") if l.startswith("{")][-1])
print("ref", r["value"], r["unit"], r["impl"], r["ms_per_step"])
PY
