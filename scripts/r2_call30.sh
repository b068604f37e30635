#!/bin/bash
mkdir -p gpurun_out
python scripts/r2_probe.py cfg5bits 2> gpurun_out/t30.err | cut -c1-420; tail -2 gpurun_out/t30.err
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "query_order" 2>&1 | tail -3
