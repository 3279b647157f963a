#!/bin/bash
set -u
mkdir -p gpurun_out
export B2S_GEMM_MC=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tc_linear or gate or resskip or fused_layer or cond_table" > gpurun_out/mc_tests.log 2>&1
echo "kernel tests rc=$?"; tail -4 gpurun_out/mc_tests.log
timeout 900 python -m pytest tests/test_gpu_tc_parity.py -m gpu -x -q -k "config5 or config3 or lynx" > gpurun_out/mc_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/mc_parity.log
for wl in config5 config3; do
for v in 1 0; do
  B2S_GEMM_MC=$v timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-secondary 2> gpurun_out/mc_bench.err | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$wl mc=$v', round(j['value']/1e6,3), round(j['ms_per_step'],2), j['clocks']['sm_mhz'], j['roofline']['frac'])"
done; done
