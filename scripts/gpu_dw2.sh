#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "dwconv" > gpurun_out/dw_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/dw_tests.log
for v in 1 0; do echo "B2S_DWCONV_MMA=$v"; B2S_DWCONV_MMA=$v timeout 300 python scripts/bench_dwconv.py 2>&1 | tail -3; done | tee gpurun_out/dw_micro.log
