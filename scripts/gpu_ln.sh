#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_parity.py tests/test_gpu_eager_baseline.py tests/test_gpu_parity.py -m gpu -x -q -k "lynx or config3" > gpurun_out/ln_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/ln_parity.log
timeout 600 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ln_bench.log 2> gpurun_out/ln_bench.err
echo "bench rc=$?"; tail -1 gpurun_out/ln_bench.log | cut -c1-200
