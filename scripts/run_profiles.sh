#!/bin/bash
# How the files under profiles/ were produced (run under gpurun on ONE B200; never wrap a multi-rank command in ncu).
# Every ncu command is preceded by the same command without ncu (it must exit 0 first).
set -eu
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
# 1. launch list of the DEFAULT bench command (CUDA-graph replay)  -> profiles/r01_launches_default_bench.csv
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 400 --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu.log 2>&1
# 2. full capture of the dominant kernel                          -> profiles/r01_ncu_full_wavenet_stack.csv
DBG="python bench.py --k-step 8 --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
$DBG > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wavenet_stack -s 4 -c 1 -o gpurun_out/prof_stack $DBG > gpurun_out/ncu2.log 2>&1
# 3. in-kernel phase timeline (needs a B2S_TLOG build)              -> profiles/r01_stack_timeline.txt
#    B2S_BUILD_TLOG=1 python xiaoicesing_io_b200/_build.py --force && python scripts/stack_timeline.py
# 4. GEMM probe grid                                                -> quoted in DESIGN.md section 3.3
#    python scripts/tc_probe.py
