#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/n8_gpus.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/n8_bench1.log 2> gpurun_out/n8_bench1.err
echo "N=1 rc=$?"; tail -1 gpurun_out/n8_bench1.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/n8_bench8.log 2> gpurun_out/n8_bench8.err
echo "N=8 rc=$?"; tail -1 gpurun_out/n8_bench8.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/n8_multi.log 2>&1
echo "multi rc=$?"; tail -2 gpurun_out/n8_multi.log
