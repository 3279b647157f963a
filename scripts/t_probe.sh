#!/bin/bash
# probe of the transposed stack kernel: both readings of the matrix-descriptor base offset for row-shifted windows
mkdir -p gpurun_out
for dbg in 0 1; do
  echo "== B2S_STACKT_DBG=$dbg"
  B2S_STACKT_DBG=$dbg timeout 300 python -m pytest tests/test_gpu_tc_parity.py -m gpu -q -x -k "transposed_stack" 2>&1 | tail -8
done
