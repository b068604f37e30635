#!/bin/bash
# ring-buffer candidate insertion (default build) against the backward shift loop (alt build): cfg3 frame + k-NN parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "knn or photon or gather or query_order or headline or cfg4" 2>&1 | tail -3
PROBE_BITS=6 PROBE_ROUNDS=8 python scripts/r2_probe.py tilesort 2> gpurun_out/t34.err | grep -o '"RT_SORT_BITS.*"shade": [0-9.]*'
RT_B200_LIB=$PWD/ray-tracing-engine_b200/lib/librt_b200_alt.so PROBE_BITS=6 PROBE_ROUNDS=8 python scripts/r2_probe.py tilesort 2>> gpurun_out/t34.err | grep -o '"RT_SORT_BITS.*"shade": [0-9.]*'
python scripts/knn_k_sweep.py 2>> gpurun_out/t34.err | tail -12
