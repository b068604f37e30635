#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
PROBE_BITS=6 PROBE_ROUNDS=8 python scripts/r2_probe.py tilesort 2> gpurun_out/t39.err | cut -c1-330
python scripts/r2_probe.py cfg4 2>> gpurun_out/t39.err | tail -1 | cut -c1-420
