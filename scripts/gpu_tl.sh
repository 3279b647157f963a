#!/bin/bash
mkdir -p gpurun_out
for B in ${TL_BS:-16 1}; do
TL_B=$B B2S_LIB=$PWD/xiaoicesing_io_b200/libb2s_tlog.so timeout 200 python scripts/stack3_timeline.py > gpurun_out/tl3_B$B.txt 2>&1; echo "tl rc=$?"; cat gpurun_out/tl3_B$B.txt
done
