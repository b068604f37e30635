#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest3.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"
python scripts/profile_frame.py cfg3 gpurun_out/frame_cfg3.json > gpurun_out/plain_cfg3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_shade -c 1 -o gpurun_out/prof_r2_knn python scripts/profile_frame.py cfg3 > gpurun_out/ncu_knn.log 2>&1; echo "ncu knn rc=$?"
python scripts/profile_frame.py cfg2 gpurun_out/frame_cfg2.json > gpurun_out/plain_cfg2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_trace|k_shade" -c 4 -o gpurun_out/prof_r2_cfg2 python scripts/profile_frame.py cfg2 > gpurun_out/ncu_cfg2full.log 2>&1; echo "ncu cfg2 rc=$?"
