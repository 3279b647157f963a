#!/bin/bash
# stack3 iteration loop: kernel test, timeline (TLOG build), short bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "stack3" > gpurun_out/t_stack3.log 2>&1; echo "test rc=$?"; tail -2 gpurun_out/t_stack3.log
timeout 600 python -m pytest tests/test_gpu_tc_parity.py -m gpu -x -q -k "stack3_matches" > gpurun_out/t_s3p.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/t_s3p.log
B2S_LIB=$PWD/xiaoicesing_io_b200/libb2s_tlog.so timeout 200 python scripts/stack3_timeline.py > gpurun_out/tl3.txt 2>&1; echo "tl rc=$?"; grep -E "^#|^    [45] |^layer|period" gpurun_out/tl3.txt | cut -c1-250
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary ${BARGS:-} > gpurun_out/ab.log 2> gpurun_out/ab.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1])
r=l["roofline"]
print("value %.3e" % l["value"], "ms/step %.2f" % l["ms_per_step"], "%.1f TF frac %.3f launch_ms %.4f" % (r["achieved"], r["frac"], r["avg_launch_ms"]), l["clocks"])
PY
