#!/bin/bash
# A/B of the generic GEMM's last-wave split (B2S_GEMM_NOSPLIT=1 turns it off): config 5, config 3, the vocoder, the aux decoder.
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for ns in 0 1; do
  if [ $ns = 1 ]; then export B2S_GEMM_NOSPLIT=1; else unset B2S_GEMM_NOSPLIT; fi
  python - <<PY
import json, torch, bench
dev = torch.device('cuda:0')
out = {}
for name in ('config5', 'config3', 'config1'):
    w = dict(bench.WORKLOADS[name])
    model = bench.make_model(w, 'fp16', dev)
    v, ms, nfe = bench.time_workload(model, w, dev, steps=3, warmup=3)
    out[name] = round(v / 1e6, 3)
    del model; torch.cuda.empty_cache()
r = bench.time_vocoder('fp16', dev)
out['voc_B1_ms'] = round(r['B1']['ms_per_call'], 3); out['voc_B8_ms'] = round(r['B8']['ms_per_call'], 3)
out['aux_ms'] = round(bench.time_aux_decoder('fp16', dev)['ms_per_call'], 4)
out['enc_ms'] = round(bench.time_acoustic_encoder('fp16', dev)['ms_per_call'], 4)
out['tok2mel_ms'] = round(bench.time_tokens_to_mel('fp16', dev)['ms_per_call'], 3)
print('nosplit=$ns', json.dumps(out), flush=True)
PY
done 2>&1 | grep -E "nosplit|rror" | tee gpurun_out/split_ab.txt
