#!/bin/bash
# How the round-2 files under profiles/ were produced (run under gpurun on ONE B200; never wrap a multi-rank command in ncu).
# Every ncu command is preceded by the same command without ncu (it must exit 0 first).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
# 1. launch list of the DEFAULT bench workload (CUDA-graph replay)   -> profiles/r02_launches_default_bench.csv
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 500 --csv --log-file gpurun_out/r02_launches_default_bench.csv $CMD > gpurun_out/ncu.log 2>&1
echo "launch list rc=$?"
# 2. full capture of the dominant kernel and of the skip / head kernel -> profiles/r02_ncu_full_*.csv
DBG="python bench.py --k-step 8 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-secondary"
$DBG > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wavenet_stack3 -s 4 -c 1 -o gpurun_out/r02_prof_stack3 -f $DBG > gpurun_out/ncu2.log 2>&1
echo "full stack3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:wavenet_skiphead3 -s 4 -c 1 -o gpurun_out/r02_prof_skiphead3 -f $DBG > gpurun_out/ncu3.log 2>&1
echo "full skiphead3 rc=$?"
ls -la gpurun_out/*.ncu-rep
