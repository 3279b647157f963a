#!/bin/bash
# Vocoder profile evidence: ncu launch list (gpu__time_duration) of ONE call at the public geometry (B = 1 and B = 8 x 690 frames,
# launched from the host, one stream), and one `--set full` capture of a late-stage conv GEMM.  Outputs under gpurun_out/.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
cat > /tmp/voc_one.py <<'PY'
import os, torch, bench
import xiaoicesing_io_b200 as P
dev = torch.device('cuda:0')
P.hparams.clear(); P.hparams.update(b2s_precision='fp16', b2s_cuda_graph=False, b2s_voc_streams=False)
torch.manual_seed(0)
gen = P.vocoder.Generator(dict(bench.VOCODER_H)).to(dev).eval()
B, T = int(os.environ.get('VB', 1)), 690
mel = torch.randn((B, T, 128), device=dev) * 1.5 - 4.0
f0 = 110.0 * 2 ** (2 * torch.rand((B, T), device=dev))
ri, nz = torch.rand(1, 1, 9, device=dev), torch.randn(B, T * 512, 9, device=dev)
for _ in range(2):
    gen.forward_rows(mel, f0, rand_ini=ri, noise=nz)
torch.cuda.synchronize()
PY
for VB in 1 8; do
  export VB
  timeout 300 python /tmp/voc_one.py || exit 1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/voc_launches_B$VB.csv python /tmp/voc_one.py > gpurun_out/voc_ncu.log 2>&1
  echo "ncu list B=$VB rc=$?"
done
# the last residual-block convs of the second call at B = 8 (stage 4): launches 2 x 109 - 12 ...
VB=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_cg2_kernel --launch-skip 190 --launch-count 2 \
  -o gpurun_out/voc_gemm_full python /tmp/voc_one.py > gpurun_out/voc_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/voc_gemm_full.ncu-rep --page raw --csv > gpurun_out/voc_gemm_full_raw.csv 2>/dev/null
ls -la gpurun_out/voc_gemm_full* | head
