#!/bin/bash
# k_shade occupancy: 8 (default build) / 9 / 10 CTAs per SM, grid = resident CTAs
mkdir -p gpurun_out
PROBE_ENV="RT_DUMMY=8" python scripts/r2_probe.py envab 2> gpurun_out/t53.err | cut -c1-330
for m in 9 10; do
  RT_B200_LIB=$PWD/ray-tracing-engine_b200/lib/librt_b200_alt$m.so PROBE_ENV="RT_DUMMY=$m" python scripts/r2_probe.py envab 2>> gpurun_out/t53.err | cut -c1-330
done
