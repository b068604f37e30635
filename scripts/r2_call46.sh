#!/bin/bash
# 16-bit stack entry distances: cfg2 frame, the 1.23 M-triangle scene, and the trace parity tests
mkdir -p gpurun_out
PROBE_ENV="RT_DUMMY=0,1" python scripts/r2_probe.py envab 2> gpurun_out/t46.err | cut -c1-330
python scripts/r2_probe.py cfg5bits 2>> gpurun_out/t46.err | head -1 | cut -c1-330
timeout 900 python -m pytest tests -x -q -m gpu -k "trace or bvh or million or headline or smoke" 2>&1 | tail -3
