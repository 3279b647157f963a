#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_dwconv.py > gpurun_out/dw_micro2.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_mma -s 3 -c 1 -o gpurun_out/dw_mma -f python scripts/bench_dwconv.py > gpurun_out/dw_ncu.log 2>&1
tail -3 gpurun_out/dw_ncu.log
