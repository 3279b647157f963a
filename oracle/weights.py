"""Oracle (TEST INFRASTRUCTURE): the seeded random-init recipe.

There is no network for checkpoints, so every parity test and bench run uses
random-init weights of the reference's architecture.  The recipe follows the
reference's initialisers in distribution (not in RNG stream):

* ``Conv1d`` subclasses in wavenet.py:12-15 / lynxnet.py:13-16 -> Kaiming normal;
* plain ``nn.Conv1d`` / ``nn.Linear``            -> U(-1/sqrt(fan_in), 1/sqrt(fan_in));
* ``LayerNorm`` weight 1 (+ small jitter so the affine is exercised), bias jitter;
* ``PReLU`` 0.25 (+ jitter).

Gotcha 1 of SURVEY.md section 8a: the reference zero-initialises the final
``output_projection.weight`` (wavenet.py:73, lynxnet.py:126), which makes the
backbone output a constant and any parity test vacuous.  The recipe re-draws it
as N(0, sigma_w^2); ``sigma_w`` defaults to 0.01 and is part of every parity
statement.
"""
from __future__ import annotations

import math

import torch

from .denoisers import LYNXNetCfg, WaveNetCfg

SIGMA_W_DEFAULT = 0.01


def _uniform(g, shape, fan_in):
    b = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape, generator=g) * 2 - 1) * b


def _kaiming(g, shape, fan_in):
    return torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)


def wavenet_state_dict(cfg: WaveNetCfg, seed: int = 0, sigma_w: float = SIGMA_W_DEFAULT):
    g = torch.Generator().manual_seed(seed)
    C, H, MF, L = cfg.num_channels, cfg.hidden_size, cfg.in_dims * cfg.n_feats, cfg.num_layers
    sd = {}
    sd['input_projection.weight'] = _kaiming(g, (C, MF, 1), MF)
    sd['input_projection.bias'] = _uniform(g, (C,), MF)
    sd['mlp.0.weight'] = _uniform(g, (4 * C, C), C)
    sd['mlp.0.bias'] = _uniform(g, (4 * C,), C)
    sd['mlp.2.weight'] = _uniform(g, (C, 4 * C), 4 * C)
    sd['mlp.2.bias'] = _uniform(g, (C,), 4 * C)
    for i in range(L):
        p = f'residual_layers.{i}.'
        sd[p + 'dilated_conv.weight'] = _uniform(g, (2 * C, C, 3), 3 * C)
        sd[p + 'dilated_conv.bias'] = _uniform(g, (2 * C,), 3 * C)
        sd[p + 'diffusion_projection.weight'] = _uniform(g, (C, C), C)
        sd[p + 'diffusion_projection.bias'] = _uniform(g, (C,), C)
        sd[p + 'conditioner_projection.weight'] = _uniform(g, (2 * C, H, 1), H)
        sd[p + 'conditioner_projection.bias'] = _uniform(g, (2 * C,), H)
        sd[p + 'output_projection.weight'] = _uniform(g, (2 * C, C, 1), C)
        sd[p + 'output_projection.bias'] = _uniform(g, (2 * C,), C)
    sd['skip_projection.weight'] = _kaiming(g, (C, C, 1), C)
    sd['skip_projection.bias'] = _uniform(g, (C,), C)
    sd['output_projection.weight'] = torch.randn((MF, C, 1), generator=g) * sigma_w
    sd['output_projection.bias'] = _uniform(g, (MF,), C)
    return sd


def lynxnet_state_dict(cfg: LYNXNetCfg, seed: int = 0, sigma_w: float = SIGMA_W_DEFAULT):
    g = torch.Generator().manual_seed(seed)
    C, H, MF, L = cfg.num_channels, cfg.hidden_size, cfg.in_dims * cfg.n_feats, cfg.num_layers
    E, K = cfg.expansion_factor, cfg.kernel_size
    inner = C * E
    sd = {}
    sd['input_projection.weight'] = _kaiming(g, (C, MF, 1), MF)
    sd['input_projection.bias'] = _uniform(g, (C,), MF)
    sd['diffusion_embedding.1.weight'] = _uniform(g, (4 * C, C), C)
    sd['diffusion_embedding.1.bias'] = _uniform(g, (4 * C,), C)
    sd['diffusion_embedding.3.weight'] = _uniform(g, (C, 4 * C), 4 * C)
    sd['diffusion_embedding.3.bias'] = _uniform(g, (C,), 4 * C)
    for i in range(L):
        p = f'residual_layers.{i}.'
        sd[p + 'diffusion_projection.weight'] = _uniform(g, (C, C, 1), C)
        sd[p + 'diffusion_projection.bias'] = _uniform(g, (C,), C)
        sd[p + 'conditioner_projection.weight'] = _uniform(g, (C, H, 1), H)
        sd[p + 'conditioner_projection.bias'] = _uniform(g, (C,), H)
        q = p + 'convmodule.'
        sd[q + 'net.0.weight'] = 1.0 + 0.1 * torch.randn((C,), generator=g)
        sd[q + 'net.0.bias'] = 0.1 * torch.randn((C,), generator=g)
        sd[q + 'net.2.weight'] = _uniform(g, (2 * inner, C, 1), C)
        sd[q + 'net.2.bias'] = _uniform(g, (2 * inner,), C)
        sd[q + 'net.4.weight'] = _uniform(g, (inner, 1, K), K)
        sd[q + 'net.4.bias'] = _uniform(g, (inner,), K)
        if cfg.activation == 'PReLU':
            sd[q + 'net.5.weight'] = 0.25 + 0.05 * torch.randn((inner,), generator=g)
        sd[q + 'net.6.weight'] = _uniform(g, (C, inner, 1), inner)
        sd[q + 'net.6.bias'] = _uniform(g, (C,), inner)
    sd['norm.weight'] = 1.0 + 0.1 * torch.randn((C,), generator=g)
    sd['norm.bias'] = 0.1 * torch.randn((C,), generator=g)
    sd['output_projection.weight'] = torch.randn((MF, C, 1), generator=g) * sigma_w
    sd['output_projection.bias'] = _uniform(g, (MF,), C)
    return sd


def make_state_dict(cfg, seed: int = 0, sigma_w: float = SIGMA_W_DEFAULT):
    if isinstance(cfg, WaveNetCfg):
        return wavenet_state_dict(cfg, seed, sigma_w)
    if isinstance(cfg, LYNXNetCfg):
        return lynxnet_state_dict(cfg, seed, sigma_w)
    raise TypeError(cfg)
