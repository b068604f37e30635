#!/bin/bash
# e2e phase breakdown of the user-facing call (cfg2)
mkdir -p gpurun_out
python scripts/r2_probe.py e2e > gpurun_out/e2e_probe.jsonl 2> gpurun_out/e2e_probe.err
RT_TIMING=1 python scripts/r2_probe.py e2e > /dev/null 2> gpurun_out/e2e_timing.err
tail -3 gpurun_out/e2e_probe.jsonl; tail -4 gpurun_out/e2e_timing.err | cut -c1-1500
