#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest8.log
python scripts/tune.py --one | sed "s/^/minb8 /"
RT_B200_LIB=$PWD/ray-tracing-engine_b200/lib/librt_b200_alt.so python scripts/tune.py --one | sed "s/^/minb7 /"
python scripts/r2_probe.py own 2>/dev/null | head -1
RT_B200_LIB=$PWD/ray-tracing-engine_b200/lib/librt_b200_alt.so python scripts/r2_probe.py own 2>/dev/null | head -1
