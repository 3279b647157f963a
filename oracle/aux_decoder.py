"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the reference's ConvNeXt aux decoder - the producer of ``x_start`` for
shallow diffusion, SURVEY.md section 8 row f-2 - as plain functions over a state dict with the reference's parameter names.

Follows, line by line:
    modules/aux_decoder/convnext.py:39-57   ConvNeXtBlock.forward   (depthwise k=7 -> LayerNorm eps 1e-6 -> Linear C->4C -> erf-GELU
                                                                     -> Linear 4C->C -> * gamma -> + residual)
    modules/aux_decoder/convnext.py:59-87   ConvNeXtDecoder         (inconv k -> blocks -> outconv k, 'same' zero padding)
    modules/aux_decoder/__init__.py:28-70   AuxDecoderAdaptor       (reshape to [B, F, T, M] for n_feats > 1, denorm_spec when infer)

Pinned by tests/golden/aux_*.npz (outputs of the unmodified reference, oracle/make_golden.py).  Only tests/, smoke() and
bench.py's CPU legs may import this module; the product never does.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class ConvNeXtCfg:
    in_dims: int = 256            # hidden_size: the decoder reads the acoustic condition [B, T, H] (toplevel.py:94)
    out_dims: int = 128           # mel bins (x num_feats)
    num_feats: int = 1
    num_channels: int = 512       # configs/acoustic.yaml:101-105
    num_layers: int = 6
    kernel_size: int = 7


def _c(sd, name, dtype):
    return sd[name].to(dtype)


def convnext_block(sd, i: int, x, dtype=torch.float32):
    """x [B, C, T] -> [B, C, T] (convnext.py:39-57)."""
    p = f'conv.{i}.'
    C = x.shape[1]
    y = F.conv1d(x, _c(sd, p + 'dwconv.weight', dtype), _c(sd, p + 'dwconv.bias', dtype), padding=3, groups=C)      # :41, k = 7 fixed (:25)
    y = y.transpose(1, 2)                                                                                          # :42
    y = F.layer_norm(y, (C,), _c(sd, p + 'norm.weight', dtype), _c(sd, p + 'norm.bias', dtype), eps=1e-6)           # :44 (:27)
    y = F.linear(y, _c(sd, p + 'pwconv1.weight', dtype), _c(sd, p + 'pwconv1.bias', dtype))                         # :45
    y = F.gelu(y)                                                                                                  # :46 exact erf GELU
    y = F.linear(y, _c(sd, p + 'pwconv2.weight', dtype), _c(sd, p + 'pwconv2.bias', dtype))                         # :47
    if p + 'gamma' in sd:
        y = _c(sd, p + 'gamma', dtype) * y                                                                         # :48-49
    return x + y.transpose(1, 2)                                                                                   # :50-54 (dropout: eval)


def convnext_decoder_forward(sd, cfg: ConvNeXtCfg, cond, dtype=torch.float32):
    """cond [B, T, in_dims] -> [B, T, out_dims * num_feats] (convnext.py:80-87)."""
    k = cfg.kernel_size
    x = cond.to(dtype).transpose(1, 2)
    x = F.conv1d(x, _c(sd, 'inconv.weight', dtype), _c(sd, 'inconv.bias', dtype), padding=(k - 1) // 2)
    for i in range(cfg.num_layers):
        x = convnext_block(sd, i, x, dtype)
    x = F.conv1d(x, _c(sd, 'outconv.weight', dtype), _c(sd, 'outconv.bias', dtype), padding=(k - 1) // 2)
    return x.transpose(1, 2)


def aux_adaptor_forward(sd, cfg: ConvNeXtCfg, cond, spec_min, spec_max, infer=True, dtype=torch.float32):
    """AuxDecoderAdaptor.forward (aux_decoder/__init__.py:56-70): [B, T, M] or [B, F, T, M], de-normalised when ``infer``.
    ``sd`` holds the decoder's parameters WITHOUT the adaptor's ``decoder.`` prefix."""
    x = convnext_decoder_forward(sd, cfg, cond, dtype)
    if cfg.num_feats > 1:
        x = x.reshape(-1, x.shape[1], cfg.num_feats, cfg.out_dims).transpose(1, 2)                                  # :61-66
    if infer:
        smin = torch.tensor(spec_min, dtype=dtype)[None, None, :].transpose(-3, -2)                                 # :43-46
        smax = torch.tensor(spec_max, dtype=dtype)[None, None, :].transpose(-3, -2)
        k, b = (smax - smin) / 2., (smax + smin) / 2.                                                               # :52-55
        x = x * k + b
    return x
