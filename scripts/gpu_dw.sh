#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "dwconv" > gpurun_out/dw_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/dw_tests.log
B2S_DWCONV_MMA=0 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "dwconv" > gpurun_out/dw_tests0.log 2>&1
echo "tests(reg) rc=$?"; tail -3 gpurun_out/dw_tests0.log
timeout 600 python -m pytest tests/test_gpu_tc_parity.py -m gpu -x -q -k "config3 or lynx" > gpurun_out/dw_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/dw_parity.log
for v in 1 0; do
  B2S_DWCONV_MMA=$v timeout 600 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/dw_bench_$v.log 2> gpurun_out/dw_bench_$v.err
  echo "bench smem=$v rc=$?"; tail -1 gpurun_out/dw_bench_$v.log | cut -c1-200
done
