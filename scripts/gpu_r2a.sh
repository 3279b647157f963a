#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_tc_parity.py tests/test_gpu_eager_baseline.py -m gpu -q -x -s > gpurun_out/t_parity16.log 2>&1; echo "parity16 rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/t_parity16.log | tail -5
echo skip-bench
exit 0
python - <<"PY"
import json
l=json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
r=l["roofline"]
print("value %.3e e2e %.3e ms/step %.2f dtype %s" % (l["value"], l["e2e"]["value"], l["ms_per_step"], l["dtype"]))
print("roofline", r["kernel"][:60], "%.1f TF frac %.3f launch_ms %.4f" % (r["achieved"], r["frac"], r["avg_launch_ms"]), "clocks", l["clocks"])
print("cpu", l["cpu_baseline"]); print("eager", l["gpu_eager_baseline"])
for k,v in (l.get("secondary") or {}).items():
    print(k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk not in ("workload",)})
PY
