#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -p no:cacheprovider -x -k "device_photon" > gpurun_out/pytest10.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest10.log
REPEAT=3 timeout 600 python scripts/dist_emulate.py 8 > gpurun_out/dist_emulate.txt 2>&1; echo "emulate rc=$?"; cat gpurun_out/dist_emulate.txt | tail -22
