#!/bin/bash
# refill threshold of the nearest-hit launches alone (any-hit stays at 14); one process per value (the constant is set once)
mkdir -p gpurun_out
for v in 14 10 18 22 26; do
  PROBE_ENV="RT_REFILL_BELOW_NEAREST=$v" python scripts/r2_probe.py envab 2>> gpurun_out/t45.err | cut -c1-330
done
