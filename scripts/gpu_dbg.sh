#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3; do
  B2S_DBG_NOLOAD=$v timeout 600 python bench.py --workload config5 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('dbg=$v', j['value']/1e6, j['ms_per_step'], j['clocks'])"
done
