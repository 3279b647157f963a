"""LYNXNet denoiser backed by libb2s (drop-in for reference modules/backbones/lynxnet.py).

Same constructor, parameter names (``residual_layers.{i}.convmodule.net.{0,2,4,5,6}``, ``norm``,
``diffusion_embedding.{1,3}``) and forward signature as the reference; arithmetic in CUDA kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..engine import LYNXNetEngine
from ..hparams import hparams
from .wavenet import Conv1d, SinusoidalPosEmb, _B2SBackbone


class _Placeholder(nn.Module):
    """Parameter-less slot keeping ``nn.Sequential`` indices identical to the reference's."""


class LYNXConvModule(nn.Module):
    """Parameter container of lynxnet.py:29-65: net.0 LayerNorm, net.2 1x1 up, net.4 depthwise,
    net.5 activation (PReLU has weights), net.6 1x1 down."""

    def __init__(self, dim, expansion_factor, kernel_size=31, activation='PReLU', dropout=0.0):
        super().__init__()
        inner_dim = dim * expansion_factor
        activation = activation if activation is not None else 'PReLU'
        if activation not in ('SiLU', 'ReLU', 'PReLU'):
            raise ValueError(f'{activation} is not a valid activation')
        act = nn.PReLU(inner_dim) if activation == 'PReLU' else _Placeholder()
        pad = kernel_size // 2
        self.net = nn.Sequential(
            nn.LayerNorm(dim),
            _Placeholder(),                                            # Transpose
            nn.Conv1d(dim, inner_dim * 2, 1),
            _Placeholder(),                                            # SwiGLU
            nn.Conv1d(inner_dim, inner_dim, kernel_size=kernel_size, padding=pad, groups=inner_dim),
            act,
            nn.Conv1d(inner_dim, dim, 1),
            _Placeholder(),                                            # Transpose
            _Placeholder(),                                            # Dropout / Identity (inference: identity)
        )


class LYNXNetResidualLayer(nn.Module):
    def __init__(self, dim_cond, dim, expansion_factor, kernel_size=31, activation='PReLU', dropout=0.0):
        super().__init__()
        self.diffusion_projection = nn.Conv1d(dim, dim, 1)
        self.conditioner_projection = nn.Conv1d(dim_cond, dim, 1)
        self.convmodule = LYNXConvModule(dim=dim, expansion_factor=expansion_factor, kernel_size=kernel_size,
                                         activation=activation, dropout=dropout)


class LYNXNet(_B2SBackbone):
    engine_cls = LYNXNetEngine

    def __init__(self, in_dims, n_feats, *, num_layers=6, num_channels=512, expansion_factor=2, kernel_size=31,
                 activation='PReLU', dropout=0.0, strong_cond=False):
        super().__init__()
        if kernel_size % 2 == 0:
            raise ValueError('even depthwise kernel sizes are not supported by the B200 path')
        self.in_dims = in_dims
        self.n_feats = n_feats
        self.num_layers = num_layers
        self.num_channels = num_channels
        self.expansion_factor = expansion_factor
        self.kernel_size = kernel_size
        self.activation = activation if activation is not None else 'PReLU'
        self.hidden_size = hparams['hidden_size']                     # lynxnet.py:113
        self.input_projection = Conv1d(in_dims * n_feats, num_channels, 1)
        self.diffusion_embedding = nn.Sequential(
            SinusoidalPosEmb(num_channels),
            nn.Linear(num_channels, num_channels * 4),
            nn.GELU(),
            nn.Linear(num_channels * 4, num_channels),
        )
        self.residual_layers = nn.ModuleList([
            LYNXNetResidualLayer(dim_cond=self.hidden_size, dim=num_channels, expansion_factor=expansion_factor,
                                 kernel_size=kernel_size, activation=activation, dropout=dropout)
            for _ in range(num_layers)
        ])
        self.norm = nn.LayerNorm(num_channels)
        self.output_projection = Conv1d(num_channels, in_dims * n_feats, kernel_size=1)
        self.strong_cond = strong_cond
        nn.init.zeros_(self.output_projection.weight)                 # lynxnet.py:126
