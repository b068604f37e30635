#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest17.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest17.log
for m in 6 0; do RT_SORT_SEGS=$m python scripts/r2_probe.py own 2>/dev/null | head -1 | sed "s/^/segs=$m /"; done
python scripts/tune.py --one
