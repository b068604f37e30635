#!/bin/bash
mkdir -p gpurun_out
for seg0 in 1 0; do
  echo "RT_SORT_SEG0=$seg0"
  RT_SORT_SEG0=$seg0 PROBE_BITS=6,7 PROBE_ROUNDS=8,16 python scripts/r2_probe.py tilesort 2> gpurun_out/t40.err | cut -c1-300
done
