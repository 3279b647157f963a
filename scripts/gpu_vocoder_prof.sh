#!/bin/bash
# ncu launch list (gpu__time_duration) of ONE vocoder call at the public geometry, B = 1 x 690 frames, launched from the host.
mkdir -p gpurun_out
cat > /tmp/voc_one.py <<'PY'
import torch, bench
import xiaoicesing_io_b200 as P
dev = torch.device('cuda:0')
P.hparams.clear(); P.hparams.update(b2s_precision='fp16', b2s_cuda_graph=False)
torch.manual_seed(0)
gen = P.vocoder.Generator(dict(bench.VOCODER_H)).to(dev).eval()
B, T = int(__import__('os').environ.get('VB', 1)), 690
mel = torch.randn((B, T, 128), device=dev) * 1.5 - 4.0
f0 = 110.0 * 2 ** (2 * torch.rand((B, T), device=dev))
ri, nz = torch.rand(1, 1, 9, device=dev), torch.randn(B, T * 512, 9, device=dev)
for _ in range(2):
    gen.forward_rows(mel, f0, rand_ini=ri, noise=nz)
torch.cuda.synchronize()
PY
export PYTHONPATH=$PWD; timeout 300 python /tmp/voc_one.py || exit 1
VB=${VB:-1} timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/voc_launches.csv python /tmp/voc_one.py > gpurun_out/voc_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/voc_launches.csv
