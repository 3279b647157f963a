#!/bin/bash
# CUDA-graph replay of the condition producers: A/B of the one-utterance tokens -> mel latency and of the two producers alone
set -u
mkdir -p gpurun_out
timeout 600 python - <<'PY'
import torch, bench
dev = torch.device('cuda:0')
for graph in (True, False):
    hp = dict(b2s_cuda_graph=graph)
    res = {n: round(f('fp16', dev, extra_hparams=hp)['ms_per_call'], 3) for n, f in (('tokens_to_mel_one_utterance', bench.time_tokens_to_mel),
                                                                                   ('aux_decoder_16x690', bench.time_aux_decoder),
                                                                                   ('acoustic_encoder_16x64', bench.time_acoustic_encoder))}
    print('b2s_cuda_graph', graph, res)
PY
