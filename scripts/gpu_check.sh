#!/bin/bash
# One gpurun call: GPU tests, smoke, short bench.  Everything logs under gpurun_out/.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu (kernels)" | tee gpurun_out/status.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s > gpurun_out/pytest_kernels.log 2>&1
rc=$?
echo "kernels rc=$rc" | tee -a gpurun_out/status.txt
tail -15 gpurun_out/pytest_kernels.log
if [ $rc -ne 0 ] && [ "${KEEP_GOING:-0}" != "1" ]; then exit $rc; fi
echo "== pytest -m gpu (parity)" | tee -a gpurun_out/status.txt
timeout 1200 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_kernels.py > gpurun_out/pytest_gpu.log 2>&1
echo "parity rc=$?" | tee -a gpurun_out/status.txt
tail -25 gpurun_out/pytest_gpu.log
echo "== smoke" | tee -a gpurun_out/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/status.txt
tail -3 gpurun_out/smoke.log
if [ "${SKIP_BENCH:-0}" = "1" ]; then exit 0; fi
echo "== bench" | tee -a gpurun_out/status.txt
timeout 900 python bench.py --steps ${BENCH_STEPS:-2} --warmup ${BENCH_WARMUP:-3} ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?" | tee -a gpurun_out/status.txt
tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
