#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "stack3" > gpurun_out/s3c_tests.log 2>&1
echo "kernel tests rc=$?"; tail -3 gpurun_out/s3c_tests.log
timeout 900 python -m pytest tests/test_gpu_tc_parity.py -m gpu -x -q -k "stack3 or narrow or config2" > gpurun_out/s3c_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/s3c_parity.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/s3c_bench.log 2> gpurun_out/s3c_bench.err
echo "bench rc=$?"; tail -1 gpurun_out/s3c_bench.log | cut -c1-250
timeout 600 python bench.py --workload config4_pitch --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/s3c_bench4.log 2> gpurun_out/s3c_bench4.err
echo "bench rc=$?"; tail -1 gpurun_out/s3c_bench4.log | cut -c1-250
