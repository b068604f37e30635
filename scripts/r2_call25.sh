#!/bin/bash
mkdir -p gpurun_out
python scripts/r2_probe.py tilesort > gpurun_out/tilesort.jsonl 2> gpurun_out/tilesort.err; cat gpurun_out/tilesort.jsonl | cut -c1-400; tail -3 gpurun_out/tilesort.err
