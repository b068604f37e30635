#!/bin/bash
# 8 GPUs: the distributed parity check (repeated: the round's buffer-lifetime race was intermittent), then bench at 8 and 4
mkdir -p gpurun_out
DIST_CHECK_REPEAT=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/dist_check.py > gpurun_out/dist50_8gpu.txt 2>&1; echo "dist rc=$?"; grep -c OK gpurun_out/dist50_8gpu.txt; grep "FAIL\|DIST_CHECK" gpurun_out/dist50_8gpu.txt | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench50_8gpu.json 2> gpurun_out/bench50.err; echo "bench8 rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench50.err | tail -5
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench50_4gpu.json 2>> gpurun_out/bench50.err; echo "bench4 rc=$?"
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench50_{n}gpu.json") if l.startswith("{")][-1])
        print(n, round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["parity"])
        for k, v in d.get("extra", {}).items():
            print("  ", k, round(v["value"]), round(v.get("e2e", {}).get("value", 0)), v.get("parity"))
    except Exception as e:
        print(n, "ERR", e)
PY
