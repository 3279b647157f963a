#!/bin/bash
# narrow-WaveNet padding: parity tests + A/B bench of config 4 (variances)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_parity.py -m gpu -x -q -s -k "narrow or variance or config4" > gpurun_out/pad_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/pad_tests.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "spec_norm" > gpurun_out/pad_norm.log 2>&1
echo "norm rc=$?"; tail -5 gpurun_out/pad_norm.log
for pad in true false; do
  timeout 600 python bench.py --workload config4 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --hparam b2s_pad_channels=$pad > gpurun_out/pad_bench_$pad.log 2> gpurun_out/pad_bench_$pad.err
  echo "bench pad=$pad rc=$?"; tail -1 gpurun_out/pad_bench_$pad.log | cut -c1-400
done
