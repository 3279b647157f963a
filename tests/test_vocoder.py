"""NSF-HiFiGAN vocoder: the oracle restatement against outputs of the unmodified reference, and host-side
checks of the product module that need no GPU (drop-in state dict incl. weight-norm checkpoints, loud failure on the CPU)."""
import pytest
import torch

import golden_util as GU
from oracle import vocoder as OV


def voc_cfg(meta) -> OV.NsfHifiGanCfg:
    h = dict(meta['h'])
    for k in ('upsample_rates', 'upsample_kernel_sizes', 'resblock_kernel_sizes'):
        h[k] = tuple(h[k])
    h['resblock_dilation_sizes'] = tuple(tuple(d) for d in h['resblock_dilation_sizes'])
    return OV.NsfHifiGanCfg(**h)


def run_oracle(fx, q=None):
    cfg = voc_cfg(fx.meta)
    kw = {} if q is None else dict(q=q)
    return OV.generator_forward(fx.sd, cfg, fx['mel'], fx['f0'], fx['rand_ini'] if 'rand_ini' in fx else None,
                                fx['noise'] if 'noise' in fx else None, **kw)


@pytest.mark.parametrize('name', GU.fixture_names('voc_'))
def test_oracle_matches_reference(name):
    """The restatement reproduces the unmodified reference's waveform (same weights, same inputs, same two random draws)."""
    fx = GU.Fixture(name)
    out = run_oracle(fx)
    ref = fx['out']
    assert out.shape == ref.shape
    err = (out - ref).abs().max().item()
    assert ref.abs().max().item() > 0.05, 'degenerate fixture'
    assert err <= 2e-6, f'{name}: oracle differs from the reference by {err:.3e}'


def test_fixture_is_sensitive_to_the_blocks():
    """A parity test on these weights is not blind: dropping one residual block moves the waveform by far more than the tolerance."""
    fx = GU.Fixture('voc_nsf_resblock1')
    sd = dict(fx.sd)
    sd['resblocks.3.convs2.1.weight'] = torch.zeros_like(sd['resblocks.3.convs2.1.weight'])
    cfg = voc_cfg(fx.meta)
    out = OV.generator_forward(sd, cfg, fx['mel'], fx['f0'], fx['rand_ini'], fx['noise'])
    assert (out - fx['out']).abs().max().item() > 1e-2


def test_operand_rounding_budget():
    """What 16-bit conv operands cost on the waveform (fp32 accumulation, fp32 residual stream): the tolerance of the GPU parity test
    (tests/test_gpu_vocoder.py) is set from this prediction."""
    fx = GU.Fixture('voc_nsf_resblock1')
    err = (run_oracle(fx, q=OV.round16(torch.float16)) - fx['out']).abs().max().item()
    assert err < 5e-3, err
