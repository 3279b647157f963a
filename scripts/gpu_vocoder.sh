#!/bin/bash
# One gpurun call for the vocoder: its GPU tests, then the bench block alone.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -s -p no:cacheprovider > gpurun_out/voc_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/voc_tests.log
grep -E "passed|failed" gpurun_out/voc_tests.log | tail -2
timeout 600 python - > gpurun_out/voc_bench.json 2> gpurun_out/voc_bench.err <<'PY'
import json, torch, bench
dev = torch.device('cuda:0')
print(json.dumps(bench.time_vocoder('fp16', dev), indent=1))
PY
echo "bench rc=$?"; cat gpurun_out/voc_bench.json | head -60
