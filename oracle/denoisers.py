"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the denoiser backbones.

Functional code over a flat ``{name: tensor}`` state dict that uses the
reference's parameter names (SURVEY.md section 8a "State-dict layout"), so a real
checkpoint slice ``diffusion.denoise_fn.*`` feeds it unchanged.  All tensors
stay on the CPU; ``dtype`` is fp32 (the reference's arithmetic) or fp64 (the
"truth" run used for tolerance budgeting).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# configs
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class WaveNetCfg:
    """Mirror of ``WaveNet.__init__`` arguments (reference wavenet.py:52-73)."""
    in_dims: int = 128          # mel bins (or repeat bins)
    n_feats: int = 1
    num_layers: int = 20
    num_channels: int = 256
    dilation_cycle_length: int = 4
    hidden_size: int = 256      # hparams['hidden_size'], wavenet.py:65

    def dilation(self, i: int) -> int:
        return 2 ** (i % self.dilation_cycle_length)   # wavenet.py:67


@dataclass(frozen=True)
class LYNXNetCfg:
    """Mirror of ``LYNXNet.__init__`` arguments (reference lynxnet.py:91-126)."""
    in_dims: int = 128
    n_feats: int = 1
    num_layers: int = 6
    num_channels: int = 512
    expansion_factor: int = 2
    kernel_size: int = 31
    activation: str = 'PReLU'
    strong_cond: bool = False
    hidden_size: int = 256      # hparams['hidden_size'], lynxnet.py:113


def _c(sd, name, dtype):
    return sd[name].to(dtype)


# --------------------------------------------------------------------------
# shared pieces
# --------------------------------------------------------------------------
def sinusoidal_pos_emb(t: torch.Tensor, dim: int, dtype) -> torch.Tensor:
    """common_layers.py:266-278.  ``t`` is [B] or [1]; returns [B or 1, dim].

    Note the ``half_dim - 1`` divisor and the sin-then-cos concatenation.  The
    reference builds the frequency vector in fp32 (``torch.arange`` ->
    ``torch.exp``) and multiplies by ``t`` (int64 or fp32) - restated as such;
    for the fp64 truth run the product is taken in fp64.
    """
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, device=t.device) * -k)            # fp32, as the reference
    if dtype == torch.float64:
        emb = t.to(torch.float64)[:, None] * freq.to(torch.float64)[None, :]
    else:
        emb = t[:, None] * freq[None, :]                   # int64*fp32 or fp32*fp32 -> fp32
    return torch.cat((emb.sin(), emb.cos()), dim=-1).to(dtype)


def mish(x):
    return x * torch.tanh(F.softplus(x))


# --------------------------------------------------------------------------
# WaveNet
# --------------------------------------------------------------------------
def wavenet_step_embedding(sd, cfg: WaveNetCfg, t, dtype=torch.float32):
    """``diffusion_embedding`` + ``mlp`` (wavenet.py:57-62, 89-90) -> [B or 1, C]."""
    e = sinusoidal_pos_emb(t, cfg.num_channels, dtype)
    e = F.linear(e, _c(sd, 'mlp.0.weight', dtype), _c(sd, 'mlp.0.bias', dtype))
    e = mish(e)
    e = F.linear(e, _c(sd, 'mlp.2.weight', dtype), _c(sd, 'mlp.2.bias', dtype))
    return e


def wavenet_residual_block(sd, cfg: WaveNetCfg, i: int, x, cond, emb, dtype=torch.float32):
    """``ResidualBlock.forward`` (wavenet.py:33-48).  x [B,C,T], cond [B,H,T], emb [B or 1, C]."""
    p = f'residual_layers.{i}.'
    C = cfg.num_channels
    d = cfg.dilation(i)
    dstep = F.linear(emb, _c(sd, p + 'diffusion_projection.weight', dtype),
                     _c(sd, p + 'diffusion_projection.bias', dtype)).unsqueeze(-1)        # :34
    c = F.conv1d(cond, _c(sd, p + 'conditioner_projection.weight', dtype),
                 _c(sd, p + 'conditioner_projection.bias', dtype))                         # :35
    y = x + dstep                                                                           # :36
    # zero padding is applied AFTER the step-embedding add
    y = F.conv1d(y, _c(sd, p + 'dilated_conv.weight', dtype), _c(sd, p + 'dilated_conv.bias', dtype),
                 padding=d, dilation=d) + c                                                 # :38
    gate, filt = y[:, :C], y[:, C:]                                                         # :41
    y = torch.sigmoid(gate) * torch.tanh(filt)                                              # :42
    y = F.conv1d(y, _c(sd, p + 'output_projection.weight', dtype),
                 _c(sd, p + 'output_projection.bias', dtype))                               # :44
    residual, skip = y[:, :C], y[:, C:]                                                     # :47
    return (x + residual) / math.sqrt(2.0), skip                                            # :48


def wavenet_forward(sd, cfg: WaveNetCfg, spec, t, cond, dtype=torch.float32, taps=None):
    """``WaveNet.forward`` (wavenet.py:75-107).

    spec [B,F,M,T], t [B] or [1] (int64 or float), cond [B,H,T] -> [B,F,M,T].
    ``taps`` (optional dict) receives intermediate tensors for unit parity tests.
    """
    spec = spec.to(dtype)
    cond = cond.to(dtype)
    B = spec.shape[0]
    x = spec.reshape(B, cfg.n_feats * cfg.in_dims, spec.shape[-1])                           # :82-85
    x = F.conv1d(x, _c(sd, 'input_projection.weight', dtype), _c(sd, 'input_projection.bias', dtype))
    x = F.relu(x)                                                                           # :86-88
    emb = wavenet_step_embedding(sd, cfg, t, dtype)                                          # :89-90
    if taps is not None:
        taps['stem'] = x
        taps['emb'] = emb
    skips = None
    for i in range(cfg.num_layers):                                                         # :92-94
        x, s = wavenet_residual_block(sd, cfg, i, x, cond, emb, dtype)
        skips = s if skips is None else skips + s
        if taps is not None:
            taps[f'x{i}'] = x
            taps[f'skip{i}'] = s
    x = skips / math.sqrt(cfg.num_layers)                                                   # :96
    x = F.conv1d(x, _c(sd, 'skip_projection.weight', dtype), _c(sd, 'skip_projection.bias', dtype))
    x = F.relu(x)                                                                           # :97-98
    x = F.conv1d(x, _c(sd, 'output_projection.weight', dtype), _c(sd, 'output_projection.bias', dtype))
    return x.reshape(B, cfg.n_feats, cfg.in_dims, x.shape[-1])                               # :100-106


# --------------------------------------------------------------------------
# LYNXNet
# --------------------------------------------------------------------------
def lynxnet_step_embedding(sd, cfg: LYNXNetCfg, t, dtype=torch.float32):
    """``diffusion_embedding`` (lynxnet.py:104-109): sinusoid -> Linear -> exact GELU -> Linear."""
    e = sinusoidal_pos_emb(t, cfg.num_channels, dtype)
    e = F.linear(e, _c(sd, 'diffusion_embedding.1.weight', dtype), _c(sd, 'diffusion_embedding.1.bias', dtype))
    e = F.gelu(e)
    e = F.linear(e, _c(sd, 'diffusion_embedding.3.weight', dtype), _c(sd, 'diffusion_embedding.3.bias', dtype))
    return e


def _lynx_activation(sd, prefix, name, x, dtype):
    if name == 'PReLU':
        return F.prelu(x, _c(sd, prefix + 'net.5.weight', dtype))
    if name == 'SiLU':
        return F.silu(x)
    if name == 'ReLU':
        return F.relu(x)
    raise ValueError(f'{name} is not a valid activation')


def lynxnet_layer(sd, cfg: LYNXNetCfg, i: int, x, cond, emb, dtype=torch.float32):
    """``LYNXNetResidualLayer.forward`` (lynxnet.py:76-87) with ``LYNXConvModule`` (:52-62).

    x [B,C,T]; cond [B,H,T]; emb [B or 1, C, 1].
    """
    p = f'residual_layers.{i}.'
    cproj = F.conv1d(cond, _c(sd, p + 'conditioner_projection.weight', dtype),
                     _c(sd, p + 'conditioner_projection.bias', dtype))
    if cfg.strong_cond:                    # front_cond_inject, :77-79
        x = x + cproj
        res = x
    else:                                  # :80-82
        res = x
        x = x + cproj
    x = x + F.conv1d(emb, _c(sd, p + 'diffusion_projection.weight', dtype),
                     _c(sd, p + 'diffusion_projection.bias', dtype))                         # :83
    q = p + 'convmodule.'
    C = cfg.num_channels
    h = F.layer_norm(x.transpose(1, 2), (C,), _c(sd, q + 'net.0.weight', dtype),
                     _c(sd, q + 'net.0.bias', dtype)).transpose(1, 2)                       # LN over channels
    h = F.conv1d(h, _c(sd, q + 'net.2.weight', dtype), _c(sd, q + 'net.2.bias', dtype))      # C -> 2*E*C
    inner = h.shape[1] // 2
    out, gate = h[:, :inner], h[:, inner:]                                                   # SwiGLU, common_layers.py:116-117
    h = out * F.silu(gate)
    pad = cfg.kernel_size // 2
    h = F.conv1d(h, _c(sd, q + 'net.4.weight', dtype), _c(sd, q + 'net.4.bias', dtype),
                 padding=pad, groups=inner)                                                  # depthwise
    h = _lynx_activation(sd, q, cfg.activation, h, dtype)
    h = F.conv1d(h, _c(sd, q + 'net.6.weight', dtype), _c(sd, q + 'net.6.bias', dtype))      # E*C -> C
    return h + res                                                                           # :86


def lynxnet_forward(sd, cfg: LYNXNetCfg, spec, t, cond, dtype=torch.float32, taps=None):
    """``LYNXNet.forward`` (lynxnet.py:128-163)."""
    spec = spec.to(dtype)
    cond = cond.to(dtype)
    B = spec.shape[0]
    x = spec.reshape(B, cfg.n_feats * cfg.in_dims, spec.shape[-1])
    x = F.conv1d(x, _c(sd, 'input_projection.weight', dtype), _c(sd, 'input_projection.bias', dtype))
    if not cfg.strong_cond:
        x = F.gelu(x)                                                                        # :142-143
    emb = lynxnet_step_embedding(sd, cfg, t, dtype).unsqueeze(-1)                             # :145
    if taps is not None:
        taps['stem'] = x
        taps['emb'] = emb
    for i in range(cfg.num_layers):                                                          # :147-148
        x = lynxnet_layer(sd, cfg, i, x, cond, emb, dtype)
        if taps is not None:
            taps[f'x{i}'] = x
    C = cfg.num_channels
    x = F.layer_norm(x.transpose(1, 2), (C,), _c(sd, 'norm.weight', dtype),
                     _c(sd, 'norm.bias', dtype)).transpose(1, 2)                            # :151
    x = F.conv1d(x, _c(sd, 'output_projection.weight', dtype), _c(sd, 'output_projection.bias', dtype))
    return x.reshape(B, cfg.n_feats, cfg.in_dims, x.shape[-1])


def make_denoiser(sd, cfg, dtype=torch.float32):
    """Returns ``fn(x[B,F,M,T], t[B or 1], cond[B,H,T]) -> [B,F,M,T]`` (seam 1 of SURVEY.md section 8b)."""
    if isinstance(cfg, WaveNetCfg):
        return lambda x, t, cond: wavenet_forward(sd, cfg, x, t, cond, dtype)
    if isinstance(cfg, LYNXNetCfg):
        return lambda x, t, cond: lynxnet_forward(sd, cfg, x, t, cond, dtype)
    raise TypeError(cfg)
