#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench9_8gpu.json 2> gpurun_out/bench9.err; echo "bench8 rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench9.err | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/dist_check.py > gpurun_out/dist9_8gpu.txt 2>&1; echo "dist rc=$?"; tail -8 gpurun_out/dist9_8gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench9_4gpu.json 2>> gpurun_out/bench9.err; echo "bench4 rc=$?"
