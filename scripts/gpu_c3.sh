#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gate or fused_layer" > gpurun_out/c3_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/c3_tests.log
timeout 900 python -m pytest tests/test_gpu_tc_parity.py -m gpu -x -q -k "config5 or config4 or narrow" > gpurun_out/c3_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/c3_parity.log
for v in 1 0; do
  B2S_GATE_CONV3=$v timeout 600 python bench.py --workload config5 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/c3_bench_$v.log 2> gpurun_out/c3_bench_$v.err
  echo "bench config5 conv3=$v rc=$?"; tail -1 gpurun_out/c3_bench_$v.log | cut -c1-200
done
