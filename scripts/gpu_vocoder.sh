#!/bin/bash
# One gpurun call for the vocoder: its GPU tests (+ the tests of every other user of the generic GEMM), then the bench
# block alone.  Outputs under gpurun_out/.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_vocoder.py -m gpu -q -s -p no:cacheprovider > gpurun_out/voc_tests.log 2>&1
echo "vocoder pytest rc=$?" | tee -a gpurun_out/voc_tests.log
grep -E "passed|failed" gpurun_out/voc_tests.log | tail -2
if [ "${VOC_ALSO:-1}" = "1" ]; then
  timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_aux_decoder.py tests/test_gpu_acoustic_encoder.py tests/test_gpu_tc_parity.py -m gpu -q -x -p no:cacheprovider > gpurun_out/voc_other_tests.log 2>&1
  echo "other pytest rc=$?"; tail -3 gpurun_out/voc_other_tests.log
fi
timeout 600 python - > gpurun_out/voc_bench.json 2> gpurun_out/voc_bench.err <<'PY'
import json, os, torch, bench
import xiaoicesing_io_b200 as P
dev = torch.device('cuda:0')
res = bench.time_vocoder('fp16', dev)
res['one_stream'] = {k: v for k, v in bench.time_vocoder('fp16', dev, extra_hparams=dict(b2s_voc_streams=False)).items() if k.startswith('B')}
print(json.dumps(res, indent=1))
PY
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/voc_bench.json'))
for k in ('B1', 'B8'):
    print(k, {a: round(b, 4) for a, b in d[k].items()}, 'one stream:', round(d['one_stream'][k]['ms_per_call'], 3))
PY
