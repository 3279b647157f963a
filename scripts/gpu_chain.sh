#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_parity.py -m gpu -x -q -k "chained or stack3_matches or narrow" > gpurun_out/chain_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/chain_tests.log
for wl in config4_pitch config4; do for v in true false; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --hparam b2s_chain_groups=$v 2> gpurun_out/chain_bench.err | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$wl chain=$v', round(j['value']/1e6,3), round(j['ms_per_step'],3), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3))"
done; done
for v in true false; do
  timeout 600 python bench.py --batch 64 --k-step 40 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --hparam b2s_chain_groups=$v 2> gpurun_out/chain_bench.err | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('config2 B=64 chain=$v', round(j['value']/1e6,3), round(j['ms_per_step'],3), j['clocks']['sm_mhz'], round(j['roofline']['frac'],3))"
done
