#!/bin/bash
mkdir -p gpurun_out
python scripts/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -25 gpurun_out/sanitize_memcheck.log
