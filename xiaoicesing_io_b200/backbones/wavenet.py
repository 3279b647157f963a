"""WaveNet denoiser backed by libb2s (drop-in for reference modules/backbones/wavenet.py).

Same constructor, same parameter names and shapes (so ``diffusion.denoise_fn.*`` checkpoints load with
``strict=True``), same ``forward(spec[B,F,M,T], diffusion_step[B or 1], cond[B,H,T]) -> [B,F,M,T]``.
The arithmetic runs in hand-written CUDA kernels through the C ABI; there is no PyTorch fallback.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import _cabi as C
from ..engine import WaveNetEngine
from ..hparams import hparams


class Conv1d(torch.nn.Conv1d):
    """Kaiming-normal initialised Conv1d (reference wavenet.py:12-15) - parameter container only."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        nn.init.kaiming_normal_(self.weight)


class SinusoidalPosEmb(nn.Module):
    """Parameter-less placeholder keeping the reference's module tree (common_layers.py:266-278);
    the embedding itself is computed by ``b2s_sinusoid_f32``."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim


class ResidualBlock(nn.Module):
    """Parameter container with the reference's names (wavenet.py:18-31).  The block's arithmetic is the
    fused gate / out kernels launched by ``WaveNetEngine``."""

    def __init__(self, encoder_hidden, residual_channels, dilation):
        super().__init__()
        self.residual_channels = residual_channels
        self.dilation = dilation
        self.dilated_conv = nn.Conv1d(residual_channels, 2 * residual_channels, kernel_size=3,
                                      padding=dilation, dilation=dilation)
        self.diffusion_projection = nn.Linear(residual_channels, residual_channels)
        self.conditioner_projection = nn.Conv1d(encoder_hidden, 2 * residual_channels, 1)
        self.output_projection = nn.Conv1d(residual_channels, 2 * residual_channels, 1)


def _time_major_cond(cond: torch.Tensor) -> torch.Tensor:
    """[B,H,T] (usually a transposed view of the encoder's [B,T,H]) -> contiguous [B,T,H]."""
    bth = cond.transpose(1, 2)
    if bth.is_contiguous():
        return bth                                  # zero-copy: the caller handed us condition.transpose(1, 2)
    B, H, T = cond.shape
    src = cond.contiguous()
    out = torch.empty((B, T, H), device=cond.device, dtype=torch.float32)
    C.transpose(src, out, B, H, T)
    return out


class _B2SBackbone(nn.Module):
    """Shared forward() plumbing of the two backbones (seam 1 of SURVEY.md section 8b)."""

    engine_cls = None

    def _engine(self):
        prec = hparams.get('b2s_precision') or os.environ.get('B2S_PRECISION', 'fp32')
        eng = self.__dict__.get('_b2s_engine')
        pad = (bool(hparams.get('b2s_pad_channels', True)), bool(hparams.get('b2s_narrow_slabs', True)))
        if eng is None or eng.precision != prec or getattr(eng, 'pad_channels', pad) != pad:
            eng = self.engine_cls(self, prec)
            eng.pad_channels = pad
            self.__dict__['_b2s_engine'] = eng
        return eng

    def invalidate(self):
        """Forces the packed (K-major, 16-bit, interleaved) weight copies and every captured CUDA graph to be rebuilt on the next
        call.  Weight changes through ``load_state_dict``, optimizer steps, ``.to()``, ``.half()`` or ``p.data = t`` are detected
        automatically (tensor version counters + storage pointers); call this after writing THROUGH a detached view that shares
        the parameter's storage (``p.data.copy_(...)`` on an EMA swap-in, custom loaders), which no counter sees."""
        eng = self.__dict__.get('_b2s_engine')
        if eng is not None:
            eng._packed_version = None
        from ..core._sampling import clear_graph_cache
        clear_graph_cache()

    @torch.no_grad()
    def forward(self, spec, diffusion_step, cond):
        """
        :param spec: [B, F, M, T]
        :param diffusion_step: [B] or [1], int64 or float
        :param cond: [B, H, T]
        :return: [B, F, M, T]
        """
        if not spec.is_cuda:
            raise C.B2SError('spec must be a CUDA tensor: this backbone has no CPU fallback')
        B, F_, M, T = spec.shape
        if B * T == 0:
            return torch.zeros_like(spec, dtype=torch.float32)
        with torch.cuda.device(spec.device):    # launches and the current stream belong to the tensors' device
            return self._forward_on_device(spec, diffusion_step, cond, B, F_, M, T)

    def _forward_on_device(self, spec, diffusion_step, cond, B, F_, M, T):
        eng = self._engine()
        cond_bth = _time_major_cond(cond.float())
        t = diffusion_step.reshape(-1).to(device=spec.device, dtype=torch.float32).contiguous()
        per_row = t.numel() > 1
        if per_row and t.numel() != B:
            raise C.B2SError(f'diffusion_step must have 1 or B={B} entries (got {t.numel()})')
        sess = eng.begin(cond_bth, t, per_row_t=per_row)
        x_bct = spec.float().reshape(B, F_ * M, T).contiguous()
        x_tm = torch.empty((B * T, F_ * M), device=spec.device)
        C.transpose(x_bct, x_tm, B, F_ * M, T)
        out_tm = torch.empty_like(x_tm)
        sess.eval(x_tm, 0, out_tm)
        out = torch.empty((B, F_ * M, T), device=spec.device)
        C.transpose(out_tm, out, B, T, F_ * M)
        return out.reshape(B, F_, M, T)


class WaveNet(_B2SBackbone):
    engine_cls = WaveNetEngine

    def __init__(self, in_dims, n_feats, *, num_layers=20, num_channels=256, dilation_cycle_length=4):
        super().__init__()
        self.in_dims = in_dims
        self.n_feats = n_feats
        self.num_layers = num_layers
        self.num_channels = num_channels
        self.hidden_size = hparams['hidden_size']                     # wavenet.py:65
        self.input_projection = Conv1d(in_dims * n_feats, num_channels, 1)
        self.diffusion_embedding = SinusoidalPosEmb(num_channels)
        self.mlp = nn.Sequential(
            nn.Linear(num_channels, num_channels * 4),
            nn.Mish(),
            nn.Linear(num_channels * 4, num_channels)
        )
        self.residual_layers = nn.ModuleList([
            ResidualBlock(encoder_hidden=self.hidden_size, residual_channels=num_channels,
                          dilation=2 ** (i % dilation_cycle_length))
            for i in range(num_layers)
        ])
        self.skip_projection = Conv1d(num_channels, num_channels, 1)
        self.output_projection = Conv1d(num_channels, in_dims * n_feats, 1)
        nn.init.zeros_(self.output_projection.weight)                 # wavenet.py:73
