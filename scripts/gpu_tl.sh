#!/bin/bash
mkdir -p gpurun_out
for D in ${TL_DBG:-0}; do
for B in ${TL_BS:-16}; do
echo "== B=$B dbg=$D"
B2S_STACK3_DBG=$D TL_B=$B B2S_LIB=$PWD/xiaoicesing_io_b200/libb2s_tlog.so timeout 200 python scripts/stack3_timeline.py > gpurun_out/tl3_B${B}_d$D.txt 2>&1; echo "tl rc=$?"; cat gpurun_out/tl3_B${B}_d$D.txt
done; done
