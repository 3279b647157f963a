#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_parity.py tests/test_gpu_kernels.py -m gpu -x -q -k "narrow or stack3 or variance or randomised or chained" > gpurun_out/narrow_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/narrow_tests.log
for v in true false; do
  timeout 600 python bench.py --workload config4 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --hparam b2s_narrow_slabs=$v 2> gpurun_out/narrow_bench.err | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('config4 narrow_slabs=$v', round(j['value']/1e6,3), round(j['ms_per_step'],3), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3))"
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary 2> gpurun_out/narrow_bench.err | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('config2', round(j['value']/1e6,3), round(j['ms_per_step'],3), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3))"
