#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest4.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest4.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench rc=$?"
RT_KNN_GATHER=0 timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench4_cfg3_inline.json 2> gpurun_out/bench4b.err; echo "bench inline rc=$?"
timeout 600 python scripts/r2_probe.py kd cfg4 > gpurun_out/probe4.jsonl 2> gpurun_out/probe4.err; echo "probe rc=$?"
