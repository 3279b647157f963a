#!/bin/bash
# launch list of the DEFAULT bench workload (CUDA-graph replay) -> profiles/r02_launches_default_bench.csv
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 500 --csv --log-file gpurun_out/r02_launches_default_bench.csv $CMD > gpurun_out/ncu.log 2>&1
echo "launch list rc=$?"
