#!/bin/bash
# One gpurun call: GPU tests, smoke, short bench.  Everything logs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" | tee gpurun_out/status.txt
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/status.txt
tail -25 gpurun_out/pytest_gpu.log
echo "== smoke" | tee -a gpurun_out/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/status.txt
tail -3 gpurun_out/smoke.log
echo "== bench" | tee -a gpurun_out/status.txt
timeout 900 python bench.py --steps ${BENCH_STEPS:-2} --warmup ${BENCH_WARMUP:-3} ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?" | tee -a gpurun_out/status.txt
tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
