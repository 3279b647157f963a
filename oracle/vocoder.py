"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the reference's NSF-HiFiGAN generator - mel + f0 -> waveform, the step AFTER the
sampling loop, SURVEY.md section 8 row f-1 - as plain functions over a state dict with the reference's parameter names AFTER
``remove_weight_norm`` (``conv_pre.weight``, ``ups.0.weight``, ``resblocks.3.convs1.0.weight``, ``noise_convs.1.bias``,
``m_source.l_linear.weight``, ``conv_post.weight`` ...).

Follows, line by line:
    modules/nsf_hifigan/models.py:36-68    ResBlock1.forward  (lrelu 0.1 -> dilated conv -> lrelu 0.1 -> conv -> + x, three times)
    modules/nsf_hifigan/models.py:79-95    ResBlock2.forward  (lrelu 0.1 -> dilated conv -> + x, twice)
    modules/nsf_hifigan/models.py:134-173  SineGen            (phase accumulation per frame, 1 + 8 harmonics, uv mix with noise)
    modules/nsf_hifigan/models.py:197-200  SourceModuleHnNSF  (tanh(linear(harmonics)))
    modules/nsf_hifigan/models.py:252-262  Generator.fastsinegen (mini_nsf)
    modules/nsf_hifigan/models.py:264-289  Generator.forward
    modules/vocoders/nsf_hifigan.py:57-69  NsfHifiGAN.spec2wav_torch (log10 -> ln mel: x 2.30259)

The two random draws of SineGen (``torch.rand(1, 1, dim)`` for the initial phases, ``torch.randn_like(sine_waves)`` for the
additive noise, models.py:147, :170) are ARGUMENTS here, so that a CPU oracle and a CUDA product can be fed the same numbers.

Pinned by tests/golden/voc_*.npz (outputs of the unmodified reference, oracle/make_golden.py).  Only tests/, smoke() and
bench.py's CPU legs may import this module; the product never does.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1                                             # models.py:15


@dataclass
class NsfHifiGanCfg:
    """The fields of the vocoder's config.json that Generator reads (models.py:206-250); defaults = the public 44.1 kHz model."""
    num_mels: int = 128
    sampling_rate: int = 44100
    upsample_rates: tuple = (8, 8, 2, 2, 2)
    upsample_kernel_sizes: tuple = (16, 16, 4, 4, 4)
    upsample_initial_channel: int = 512
    resblock: str = '1'
    resblock_kernel_sizes: tuple = (3, 7, 11)
    resblock_dilation_sizes: tuple = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    mini_nsf: bool = False
    harmonic_num: int = 8                                     # models.py:222
    sine_amp: float = 0.1                                     # models.py:135-136 defaults
    noise_std: float = 0.003
    voiced_threshold: float = 0.0
    extra: dict = field(default_factory=dict)

    @property
    def hop(self) -> int:
        return int(np.prod(self.upsample_rates))


def _get_padding(kernel_size, dilation=1):                    # modules/nsf_hifigan/utils.py:12-13
    return int((kernel_size * dilation - dilation) / 2)


def _ident(t):
    return t


def _conv(sd, name, x, *, dilation=1, padding=0, stride=1, q=_ident):
    return F.conv1d(q(x), q(sd[name + '.weight']), sd[name + '.bias'], stride=stride, padding=padding, dilation=dilation)


def resblock1(sd, prefix: str, x, kernel_size: int, dilations, q=_ident):
    """models.py:60-68."""
    for m, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = _conv(sd, f'{prefix}convs1.{m}', xt, dilation=d, padding=_get_padding(kernel_size, d), q=q)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = _conv(sd, f'{prefix}convs2.{m}', xt, dilation=1, padding=_get_padding(kernel_size, 1), q=q)
        x = xt + x
    return x


def resblock2(sd, prefix: str, x, kernel_size: int, dilations, q=_ident):
    """models.py:90-95."""
    for m, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = _conv(sd, f'{prefix}convs.{m}', xt, dilation=d, padding=_get_padding(kernel_size, d), q=q)
        x = xt + x
    return x


def sine_gen(cfg: NsfHifiGanCfg, f0, upp: int, rand_ini, noise):
    """SineGen.forward (models.py:134-173).  f0 [B, T]; rand_ini [dim] (entry 0 is forced to 0 like :148); noise [B, T * upp, dim]
    standard normal.  Returns [B, T * upp, dim]."""
    dim = cfg.harmonic_num + 1
    f0 = f0.unsqueeze(-1)                                                                   # :160
    rad = f0 / cfg.sampling_rate * torch.arange(1, upp + 1, device=f0.device)               # :138
    rad2 = torch.fmod(rad[..., -1:].float() + 0.5, 1.0) - 0.5                               # :139
    rad_acc = rad2.cumsum(dim=1).fmod(1.0).to(f0)                                           # :140
    rad = rad + F.pad(rad_acc[:, :-1, :], (0, 0, 1, 0))                                     # :141
    rad = rad.reshape(f0.shape[0], -1, 1)                                                   # :142
    rad = torch.multiply(rad, torch.arange(1, dim + 1, device=f0.device).reshape(1, 1, -1))  # :143
    ini = rand_ini.reshape(1, 1, dim).clone().to(rad)
    ini[..., 0] = 0                                                                          # :145-146
    rad = rad + ini
    sines = torch.sin(2 * np.pi * rad)                                                      # :147
    sine_waves = sines * cfg.sine_amp                                                       # :161
    uv = (f0 > cfg.voiced_threshold).float()                                                # :162
    uv = F.interpolate(uv.transpose(2, 1), scale_factor=upp, mode='nearest').transpose(2, 1)  # :163
    noise_amp = uv * cfg.noise_std + (1 - uv) * cfg.sine_amp / 3                            # :164
    return sine_waves * uv + noise_amp * noise                                              # :165-166


def source_module(sd, cfg: NsfHifiGanCfg, f0, upp: int, rand_ini, noise):
    """SourceModuleHnNSF.forward (models.py:197-200) -> [B, T * upp, 1]."""
    sw = sine_gen(cfg, f0, upp, rand_ini, noise)
    return torch.tanh(F.linear(sw, sd['m_source.l_linear.weight'], sd['m_source.l_linear.bias']))


def fast_sine_gen(cfg: NsfHifiGanCfg, f0):
    """Generator.fastsinegen (models.py:252-262), mini_nsf models -> [B, 1, T * upp]."""
    source_sr = cfg.sampling_rate / int(np.prod(cfg.upsample_rates[2:]))
    upp = int(np.prod(cfg.upsample_rates[:2]))
    n = torch.arange(1, upp + 1, device=f0.device)
    s0 = f0.unsqueeze(-1) / source_sr
    ds0 = F.pad(s0[:, 1:, :] - s0[:, :-1, :], (0, 0, 0, 1))
    rad = s0 * n + 0.5 * ds0 * n * (n - 1) / upp
    rad2 = torch.fmod(rad[..., -1:].float() + 0.5, 1.0) - 0.5
    rad_acc = rad2.cumsum(dim=1).fmod(1.0).to(f0)
    rad = rad + F.pad(rad_acc[:, :-1, :], (0, 0, 1, 0))
    rad = rad.reshape(f0.shape[0], 1, -1)
    return torch.sin(2 * np.pi * rad)


def generator_forward(sd, cfg: NsfHifiGanCfg, x, f0, rand_ini=None, noise=None, q=_ident, taps=None):
    """Generator.forward (models.py:264-289): x [B, num_mels, T] ln-mel, f0 [B, T] Hz -> [B, 1, T * hop] in (-1, 1).
    ``q`` rounds the operands of every dense convolution (identity for the fp32 reference; a 16-bit round trip to predict what the
    tensor-core product may differ by).  ``taps`` (a dict) receives the intermediate tensors of each stage for debugging."""
    nk = len(cfg.resblock_kernel_sizes)
    if cfg.mini_nsf:
        har = fast_sine_gen(cfg, f0)                                                        # :265-266
    else:
        har = source_module(sd, cfg, f0, cfg.hop, rand_ini, noise).transpose(1, 2)          # :267-268
    if taps is not None:
        taps['har'] = har
    x = _conv(sd, 'conv_pre', x, padding=3, q=q)                                            # :269
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)                                                    # :271
        x = F.conv_transpose1d(q(x), q(sd[f'ups.{i}.weight']), sd[f'ups.{i}.bias'], stride=u, padding=(k - u) // 2)   # :272
        if not cfg.mini_nsf:
            if i + 1 < len(cfg.upsample_rates):                                             # :238-243
                s = int(np.prod(cfg.upsample_rates[i + 1:]))
                xs_ = F.conv1d(har, sd[f'noise_convs.{i}.weight'], sd[f'noise_convs.{i}.bias'], stride=s, padding=s // 2)
            else:
                xs_ = F.conv1d(har, sd[f'noise_convs.{i}.weight'], sd[f'noise_convs.{i}.bias'])
            x = x + xs_                                                                     # :273-275
        elif i == 1:
            x = x + F.conv1d(har, sd['source_conv.weight'], sd['source_conv.bias'])         # :276-278
        if taps is not None:
            taps[f'stage{i}_in'] = x
        xs = None
        for j in range(nk):                                                                 # :279-284
            kj, dj = cfg.resblock_kernel_sizes[j], cfg.resblock_dilation_sizes[j]
            fn = resblock1 if cfg.resblock == '1' else resblock2
            r = fn(sd, f'resblocks.{i * nk + j}.', x, kj, dj, q=q)
            xs = r if xs is None else xs + r
        x = xs / nk                                                                         # :285
        if taps is not None:
            taps[f'stage{i}_out'] = x
    x = F.leaky_relu(x)                                                                     # :286 (default slope 0.01)
    x = _conv(sd, 'conv_post', x, padding=3)                                                # :287
    return torch.tanh(x)                                                                    # :288


def spec2wav(sd, cfg: NsfHifiGanCfg, mel, f0, mel_base='e', rand_ini=None, noise=None, q=_ident):
    """NsfHifiGAN.spec2wav_torch (vocoders/nsf_hifigan.py:57-69): mel [B, T, bins] -> flat waveform [B * T * hop]."""
    c = mel.transpose(2, 1)
    if mel_base != 'e':
        assert mel_base in [10, '10'], "mel_base must be 'e', '10' or 10."
        c = 2.30259 * c
    return generator_forward(sd, cfg, c, f0, rand_ini, noise, q=q).view(-1)


def round16(dtype):
    """An operand-rounding function for ``generator_forward(q=...)``."""
    return lambda t: t.to(dtype).to(torch.float32)


def random_state_dict(cfg: NsfHifiGanCfg, seed: int, gain: float = 1.0):
    """A random state dict with the reference's names and shapes whose activations stay O(1) through the network (the reference's
    own init draws N(0, 0.01) weights, models.py:44,55,249-250, which would make every residual block a near-identity and a
    parity test blind): fan-in scaled normal weights, small biases.  Used by tests, smoke() and bench.py."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def put(name, shape, fan_in, b_shape, scale=1.0):
        sd[name + '.weight'] = torch.randn(shape, generator=g) * (gain * scale / np.sqrt(fan_in))
        sd[name + '.bias'] = torch.randn(b_shape, generator=g) * 0.05

    ch = cfg.upsample_initial_channel
    put('conv_pre', (ch, cfg.num_mels, 7), cfg.num_mels * 7, (ch,))
    nk = len(cfg.resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        put(f'ups.{i}', (ch, ch // 2, k), ch * k / u, (ch // 2,))
        ch //= 2
        for j in range(nk):
            kj, dj = cfg.resblock_kernel_sizes[j], cfg.resblock_dilation_sizes[j]
            for m in range(len(dj)):
                if cfg.resblock == '1':
                    put(f'resblocks.{i * nk + j}.convs1.{m}', (ch, ch, kj), ch * kj, (ch,))
                    put(f'resblocks.{i * nk + j}.convs2.{m}', (ch, ch, kj), ch * kj, (ch,), scale=0.5)
                else:
                    put(f'resblocks.{i * nk + j}.convs.{m}', (ch, ch, kj), ch * kj, (ch,), scale=0.5)
        if not cfg.mini_nsf:
            if i + 1 < len(cfg.upsample_rates):
                s = int(np.prod(cfg.upsample_rates[i + 1:]))
                put(f'noise_convs.{i}', (ch, 1, 2 * s), 2 * s, (ch,))
            else:
                put(f'noise_convs.{i}', (ch, 1, 1), 1, (ch,))
        elif i == 1:
            put('source_conv', (ch, 1, 1), 1, (ch,))
    if not cfg.mini_nsf:
        put('m_source.l_linear', (1, cfg.harmonic_num + 1), cfg.harmonic_num + 1, (1,), scale=3.0)
    put('conv_post', (1, ch, 7), ch * 7, (1,))
    return sd
