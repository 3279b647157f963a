"""Full-size parity + "eager PyTorch on the same B200" baseline (*that, not the CPU, is the number to
beat*).  The oracle's denoiser is plain torch functional code over a state dict, so the same restatement runs on the GPU
when its tensors live there; this file (tests/ may use the oracle) evaluates ONE denoiser call at BASELINE config 2's full
batch (B = 16 utterances x T = 690 frames, WaveNet 20 x 256) three ways:

* oracle on the GPU in strict fp32 (TF32 off)          -> the truth for a FULL-SIZE parity check of the product path
  (fp32 product <= 1e-3, bf16 / fp16 product <= 2e-2 max-abs on the denoiser output);
* oracle on the GPU the way the reference would run it  -> eager fp32 (torch defaults) and ``torch.autocast(bf16)``, timed with
  CUDA events; the product's bf16 denoiser call is timed next to them and must be faster.

The timings are written to ``gpurun_out/eager_gpu_baseline.json`` (informational; bench.py never imports the oracle for this)."""
from __future__ import annotations

import json
import os

import pytest
import torch

from oracle import denoisers as OD
from oracle import weights as OW

pytestmark = pytest.mark.gpu

B, T, M, H = 16, 690, 128, 256


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    return torch.device('cuda', 0)


def _inputs(dev):
    g = torch.Generator().manual_seed(11)
    spec = torch.randn((B, 1, M, T), generator=g).to(dev)
    cond = torch.randn((B, H, T), generator=g).to(dev)
    t = torch.full((B,), 217, dtype=torch.long, device=dev)
    return spec, t, cond


def _product(precision, sd, dev):
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    P.hparams.update(hidden_size=H, b2s_precision=precision)
    net = P.build_backbone(M, 1, 'wavenet', dict(num_layers=20, num_channels=256, dilation_cycle_length=4))
    net.load_state_dict(sd, strict=True)
    return net.to(dev).eval()


def _time(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def test_full_size_parity_and_eager_baseline(dev):
    cfg = OD.WaveNetCfg()
    sd = OW.make_state_dict(cfg, seed=0, sigma_w=0.01)
    sd_gpu = {k: v.to(dev) for k, v in sd.items()}
    spec, t, cond = _inputs(dev)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        with torch.no_grad():
            truth = OD.wavenet_forward(sd_gpu, cfg, spec, t, cond)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    scale = float(truth.abs().max())
    report = dict(workload='config2 denoiser call: WaveNet 20x256, B=16, T=690', truth_absmax=scale)
    for precision, tol in (('fp32', 1e-3), ('bf16', 2e-2), ('fp16', 2e-2)):
        net = _product(precision, sd, dev)
        with torch.no_grad():
            out = net(spec, t, cond)
        err = float((out - truth).abs().max())
        report[f'max_abs_{precision}'] = err
        assert out.shape == truth.shape
        assert err <= tol, (precision, err, scale)

    def eager():
        with torch.no_grad():
            return OD.wavenet_forward(sd_gpu, cfg, spec, t, cond)

    def eager_bf16():
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
            return OD.wavenet_forward(sd_gpu, cfg, spec, t, cond)

    net = _product('bf16', sd, dev)

    def product():
        with torch.no_grad():
            return net(spec, t, cond)

    ms = dict(eager_fp32=_time(eager), eager_autocast_bf16=_time(eager_bf16), product_bf16_call=_time(product))
    for k, v in ms.items():
        report[f'ms_{k}'] = v
        report[f'frames_per_s_{k}'] = B * T / (v * 1e-3)
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/eager_gpu_baseline.json', 'w') as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report))
    assert ms['product_bf16_call'] < ms['eager_autocast_bf16'], ms
