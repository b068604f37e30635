#!/bin/bash
# 2 GPUs: bench at 2 and 1 (same box), the 2-GPU distributed test, the new BVH builder test
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench51_2gpu.json 2> gpurun_out/bench51.err; echo "bench2 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench51_1gpu.json 2>> gpurun_out/bench51.err; echo "bench1 rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench51_ref.json 2>> gpurun_out/bench51.err; echo "ref rc=$?"
python - <<'PY'
import json
for n in (2, 1):
    d = json.loads([l for l in open(f"gpurun_out/bench51_{n}gpu.json") if l.startswith("{")][-1])
    print(n, round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d.get("cpu_baseline"))
    for k, v in d.get("extra", {}).items():
        if "value" in v: print("  ", k, round(v["value"]), round(v.get("e2e", {}).get("value", 0)), v["roofline"].get("frac"))
        else: print("  ", k, v.get("ms_per_call"), v.get("speedup_vs_reference_program"))
print(open("gpurun_out/bench51_ref.json").read()[-600:])
PY
