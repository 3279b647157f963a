#!/bin/bash
# stack3 iteration loop: kernel test, timeline (TLOG build), short bench A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "stack3" > gpurun_out/t_stack3.log 2>&1; echo "test rc=$?"; tail -3 gpurun_out/t_stack3.log
B2S_STACK3_CLUSTER=2 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "stack3" > gpurun_out/t_stack3_c2.log 2>&1; echo "test(cluster 2) rc=$?"; tail -3 gpurun_out/t_stack3_c2.log
B2S_LIB=$PWD/xiaoicesing_io_b200/libb2s_tlog.so timeout 200 python scripts/stack3_timeline.py > gpurun_out/tl3.txt 2>&1; echo "tl rc=$?"; cat gpurun_out/tl3.txt
for v in ${AB:-true}; do
python bench.py --k-step 40 --steps 2 --warmup 3 --no-cpu-baseline --hparam b2s_stack3=$v ${BARGS:-} > gpurun_out/ab_$v.log 2> gpurun_out/ab_$v.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open("gpurun_out/ab_$v.log").read().strip().splitlines()[-1])
r=l["roofline"]
print("$v", "value %.3e" % l["value"], "ms/step %.2f" % l["ms_per_step"], r["kernel"][:30], "%.1f TF frac %.3f launch_ms %.4f" % (r["achieved"], r["frac"], r["avg_launch_ms"]), l["clocks"])
PY
done
