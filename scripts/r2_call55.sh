#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  echo "RT_L2_PERSIST=$v"
  RT_L2_PERSIST=$v RT_TIMING=1 python scripts/r2_probe.py cfg5bits 2> gpurun_out/t55_$v.err | head -1 | cut -c1-330
  grep "L2 persisting" gpurun_out/t55_$v.err | head -1
done
