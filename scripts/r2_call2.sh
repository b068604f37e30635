#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke2.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest2.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"
timeout 600 python scripts/r2_probe.py batch cfg4 > gpurun_out/probe2.jsonl 2> gpurun_out/probe2.err; echo "probe rc=$?"
