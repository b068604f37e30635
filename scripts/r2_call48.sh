#!/bin/bash
# 40-register / 12-CTA trace kernels (alt build) on the L2-latency-bound 1.23 M-triangle scene and on cfg2
mkdir -p gpurun_out
export RT_B200_LIB=$PWD/ray-tracing-engine_b200/lib/librt_b200_alt.so
python scripts/r2_probe.py cfg5bits 2> gpurun_out/t48.err | head -1 | cut -c1-330
PROBE_ENV="RT_DUMMY=0" python scripts/r2_probe.py envab 2>> gpurun_out/t48.err | cut -c1-330
