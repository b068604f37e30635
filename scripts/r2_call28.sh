#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
PROBE_BITS=6,7 PROBE_ROUNDS=8 python scripts/r2_probe.py tilesort 2> gpurun_out/t28.err | grep -o '"RT_SORT_BITS.*"shade": [0-9.]*'
python scripts/r2_probe.py cfg4 2>> gpurun_out/t28.err | cut -c1-600
