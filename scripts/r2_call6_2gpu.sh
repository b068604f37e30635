#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_edges.py -m gpu -q --tb=short -p no:cacheprovider -k "two_gpu" > gpurun_out/pytest6.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest6.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/bench6_2gpu.json 2> gpurun_out/bench6.err; echo "bench2 rc=$?"; tail -3 gpurun_out/bench6.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench6_1gpu.json 2>> gpurun_out/bench6.err; echo "bench1 rc=$?"
