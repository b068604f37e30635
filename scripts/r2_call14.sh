#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest14.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest14.log
python scripts/tune.py --one
python scripts/r2_probe.py own 2>/dev/null | head -1
RT_SCENE=stock python scripts/tune.py --one
