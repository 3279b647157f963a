#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_parity.py -m gpu -x -q -s -k "stack3_matches" > gpurun_out/t_s3p.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/t_s3p.log
for v in ${AB:-true false}; do
python bench.py --k-step 40 --steps 2 --warmup 3 --no-cpu-baseline --hparam b2s_stack3_head=$v ${BARGS:-} > gpurun_out/abh_$v.log 2> gpurun_out/abh_$v.err; echo "bench rc=$?"; tail -3 gpurun_out/abh_$v.err
python - <<PY
import json
l=json.loads(open("gpurun_out/abh_$v.log").read().strip().splitlines()[-1])
r=l["roofline"]
print("head3=$v", "value %.3e" % l["value"], "ms/step %.2f" % l["ms_per_step"], r["kernel"][:30], "%.1f TF frac %.3f launch_ms %.4f" % (r["achieved"], r["frac"], r["avg_launch_ms"]), l["clocks"])
PY
done
