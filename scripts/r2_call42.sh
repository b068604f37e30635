#!/bin/bash
mkdir -p gpurun_out
PROBE_ENV="RT_BOUNCE_OCTANT=0,1,0,1" python scripts/r2_probe.py envab 2> gpurun_out/t42.err | cut -c1-330; tail -3 gpurun_out/t42.err
