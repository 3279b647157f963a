#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_segments.py tests/test_gpu_multi.py -m gpu -q -x -s > gpurun_out/t_seg.log 2>&1; echo "segments+multi rc=$?"; grep -E "passed|failed|FAILED|Error|two-rank" gpurun_out/t_seg.log | tail -8
