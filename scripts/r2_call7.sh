#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest7.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest7.log
for rb in 10 12 14 16 18 20 22; do
  RT_REFILL_BELOW=$rb python scripts/tune.py --one 2>/dev/null | sed "s/^/refill=$rb /"
done > gpurun_out/tune7.txt
cat gpurun_out/tune7.txt
RT_SCENE=stock python scripts/tune.py --one
