"""GPU parity of the ConvNeXt aux decoder (producer of x_start) against the CPU oracle
(oracle/aux_decoder.py, pinned by tests/golden/aux_*.npz) through the product's public API.

Tolerance: the de-normalised mel within 2e-2 ABSOLUTE for fp16 operands (north_star's 16-bit bound; x_start is afterwards noised
to t = K_step, so this is far tighter than the path needs); bf16 reported with the regression guard of tests/tol16.py."""
import pytest
import torch

from oracle import aux_decoder as OA
from tol16 import check16

pytestmark = pytest.mark.gpu


def _random_sd(cfg: OA.ConvNeXtCfg, seed):
    import xiaoicesing_io_b200 as P
    torch.manual_seed(seed)
    m = P.AuxDecoderAdaptor(in_dims=cfg.in_dims, out_dims=cfg.out_dims, num_feats=cfg.num_feats, spec_min=None, spec_max=None,
                            aux_decoder_arch='convnext',
                            aux_decoder_args=dict(num_channels=cfg.num_channels, num_layers=cfg.num_layers, kernel_size=cfg.kernel_size))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith('gamma'):
                p.copy_(0.2 + 0.3 * torch.rand(p.shape, generator=g))          # trained layer scales are O(0.1 .. 1), the init 1e-6
            elif n.endswith('norm.weight') or n.endswith('norm.bias'):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return {k: v.detach().clone() for k, v in m.decoder.state_dict().items()}


@pytest.mark.parametrize('precision', ['fp16', 'bf16'])
@pytest.mark.parametrize('case', ['acoustic_default', 'feats2_k5', 'ragged_tiles', 'train_mode'])
def test_aux_decoder_against_oracle(case, precision):
    import xiaoicesing_io_b200 as P
    dev = torch.device('cuda:0')
    if case == 'acoustic_default':        # configs/acoustic.yaml:101-105: 512 channels, 6 layers, k = 7, 128 mel bins, H = 256
        cfg, B, T, smin, smax, infer = OA.ConvNeXtCfg(), 2, 300, [-12.] * 128, [0.] * 128, True
    elif case == 'feats2_k5':
        cfg = OA.ConvNeXtCfg(in_dims=64, out_dims=8, num_feats=2, num_channels=128, num_layers=3, kernel_size=5)
        B, T, smin, smax, infer = 3, 131, [[-10.] * 8, [-4.] * 8], [[2.] * 8, [6.] * 8], True
    elif case == 'ragged_tiles':          # utterances shorter than the kernel / not a multiple of any tile; 1 frame
        cfg = OA.ConvNeXtCfg(in_dims=64, out_dims=16, num_feats=1, num_channels=128, num_layers=2, kernel_size=7)
        B, T, smin, smax, infer = 5, 3, [-12.] * 16, [0.] * 16, True
    else:
        cfg, B, T, smin, smax, infer = OA.ConvNeXtCfg(num_channels=256, num_layers=2), 1, 257, [-12.] * 128, [0.] * 128, False
    sd = _random_sd(cfg, 5)
    P.hparams.clear()
    P.hparams.update(b2s_precision=precision)
    model = P.AuxDecoderAdaptor(in_dims=cfg.in_dims, out_dims=cfg.out_dims, num_feats=cfg.num_feats, spec_min=smin, spec_max=smax,
                                aux_decoder_arch='convnext',
                                aux_decoder_args=dict(num_channels=cfg.num_channels, num_layers=cfg.num_layers, kernel_size=cfg.kernel_size))
    model.decoder.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(9)
    cond = torch.randn((B, T, cfg.in_dims), generator=g)
    out = model(cond.to(dev), infer=infer)
    ref = OA.aux_adaptor_forward(sd, cfg, cond, smin, smax, infer=infer, dtype=torch.float64)
    assert out.shape == ref.shape
    assert bool(torch.isfinite(out).all())
    err, scale = float((out.double().cpu() - ref).abs().max()), float(ref.abs().max())
    print(dict(test='aux_decoder', case=case, precision=precision, max_abs=err, ref_absmax=scale))
    check16(precision, err, scale)
    # a second call re-uses the packed weights; a weight update is picked up
    out2 = model(cond.to(dev), infer=infer)
    assert torch.equal(out, out2)
    with torch.no_grad():
        model.decoder.outconv.bias.add_(1.0)
    out3 = model(cond.to(dev), infer=infer)
    k = 1.0 if not infer else float((torch.tensor(smax).reshape(-1)[0] - torch.tensor(smin).reshape(-1)[0]) / 2)
    assert abs(float((out3 - out).mean()) - k) < 1e-3 * max(1.0, k) or cfg.num_feats > 1


def test_empty_batch():
    import xiaoicesing_io_b200 as P
    P.hparams.clear()
    dec = P.build_aux_decoder(64, 16, 'convnext', dict(num_channels=128, num_layers=1)).cuda()
    assert dec(torch.zeros(0, 10, 64, device='cuda')).shape == (0, 10, 16)
    assert dec(torch.zeros(2, 0, 64, device='cuda')).shape == (2, 0, 16)
