#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest15.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest15.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench15.json 2> gpurun_out/bench15.err; echo "bench rc=$?"; tail -3 gpurun_out/bench15.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench15_ref.json 2>> gpurun_out/bench15.err; echo "ref rc=$?"
