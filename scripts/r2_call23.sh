#!/bin/bash
mkdir -p gpurun_out
python scripts/r2_probe.py sortbits > gpurun_out/sortbits.jsonl 2> gpurun_out/sortbits.err; cat gpurun_out/sortbits.jsonl | cut -c1-420; tail -3 gpurun_out/sortbits.err
