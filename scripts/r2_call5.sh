#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest5.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest5.log
B="python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $B > gpurun_out/b5_default.json 2>gpurun_out/b5.err; echo rc=$?
RT_SORT_SEG0=0 timeout 300 $B > gpurun_out/b5_seg0_unsorted.json 2>>gpurun_out/b5.err; echo rc=$?
RT_B200_LIB=$PWD/ray-tracing-engine_b200/lib/librt_b200_alt.so timeout 300 $B > gpurun_out/b5_frontscan.json 2>>gpurun_out/b5.err; echo rc=$?
RT_KNN_GATHER=1 RT_SORT_SEG0=0 timeout 300 $B > gpurun_out/b5_gather_seg0_unsorted.json 2>>gpurun_out/b5.err; echo rc=$?
RT_SORT_HITS=0 timeout 300 $B > gpurun_out/b5_nosort.json 2>>gpurun_out/b5.err; echo rc=$?
