#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_parity.py tests/test_gpu_kernels.py tests/test_gpu_segments.py -m gpu -x -q -k "narrow or stack3 or randomised or chained or segments or ragged" > gpurun_out/narrow_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/narrow_tests.log
TL_PREC=fp16 bash scripts/gpu_tl.sh | sed -n 4,9p | cut -c1-250
for rep in 1 2; do for lib in libb2s.so libb2s_prev.so; do
  B2S_LIB=$PWD/xiaoicesing_io_b200/$lib timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary 2> gpurun_out/ab.err | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('config2 $lib', round(j['value']/1e6,3), round(j['ms_per_step'],3), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3), round(j['roofline']['avg_launch_ms'],4))"
done; done
for lib in libb2s.so libb2s_prev.so; do
  B2S_LIB=$PWD/xiaoicesing_io_b200/$lib timeout 600 python bench.py --workload config4 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary 2> gpurun_out/ab.err | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('config4 $lib', round(j['value']/1e6,3), round(j['ms_per_step'],3), j['clocks']['sm_mhz'])"
done
