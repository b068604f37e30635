#!/bin/bash
# CTA-per-mesh BVH build: parity tests on both device paths, build times, e2e phases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bvh or smoke or headline" 2>&1 | tail -5
python scripts/r2_probe.py bvh e2e > gpurun_out/bvh_probe.jsonl 2> gpurun_out/bvh_probe.err; cat gpurun_out/bvh_probe.jsonl; tail -3 gpurun_out/bvh_probe.err
