#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_segments.py -m gpu -x -q > gpurun_out/tab_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/tab_parity.log
timeout 900 python -m pytest tests/test_gpu_tc_parity.py -m gpu -x -q -k "graph or config1 or weights or stale or invalidate" > gpurun_out/tab_parity2.log 2>&1
echo "parity2 rc=$?"; tail -3 gpurun_out/tab_parity2.log
for wl in config1 config4; do
timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2> gpurun_out/tab_bench.err | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$wl', round(j['value']/1e6,3), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']/1e6,3), j['gpu_launches'])"
done
