"""RectifiedFlow and its repeat-bin variants on the B200 path.

Drop-in for the reference's modules/core/reflow.py:13-261 (class names, constructor signatures,
non-persistent ``spec_min/max`` buffers, hparams keys T_start_infer / sampling_algorithm /
sampling_steps, error behaviour).  Euler / RK2 / RK4 / RK5 integration is compiled by ``schedules``.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .. import schedules as S
from ..backbones import build_backbone
from ..hparams import hparams
from ._sampling import SamplerBase
from ._variants import MultiCurveMixin, PitchClipMixin, RepeatBinsMixin


class RectifiedFlow(SamplerBase):
    backbone_attr = 'velocity_fn'

    def __init__(self, out_dims, num_feats=1, t_start=0., time_scale_factor=1000,
                 backbone_type=None, backbone_args=None,
                 spec_min=None, spec_max=None):
        super().__init__()
        self.velocity_fn = build_backbone(out_dims, num_feats, backbone_type, backbone_args)
        self.use_shallow_diffusion = hparams.get('use_shallow_diffusion', False)
        if self.use_shallow_diffusion:
            assert 0. <= t_start <= 1., 'T_start should be in [0, 1].'
        else:
            t_start = 0.
        self.t_start = t_start
        self.time_scale_factor = time_scale_factor
        self._init_common(out_dims, num_feats, spec_min, spec_max, persistent_bounds=False)   # reflow.py:33-34

    def build_program(self) -> S.Program:
        return S.build_reflow(hparams['sampling_algorithm'], hparams['sampling_steps'],
                              hparams.get('T_start_infer', self.t_start), self.use_shallow_diffusion,
                              self.time_scale_factor)

    @torch.no_grad()
    def inference(self, cond, b=1, x_end=None, device=None, lengths=None, initial_noise=None):
        """cond [B, H, T]; x_end normalised [B, F, M, T] or None  ->  [B, T, M] / [B, F, T, M]."""
        return self._run(cond, b, x_end, device, lengths, initial_noise)

    @torch.no_grad()
    def _training_forward(self, spec, cond, b, device):
        """``forward(infer=False)`` (reflow.py:36-54): v_pred at x_t = x0 + t (x1 - x0), t ~ U[t_start, 1].
        Inference kernels only - validation, not backprop."""
        t = self.t_start + (1.0 - self.t_start) * torch.rand((b,), device=device)
        x0 = torch.randn_like(spec)
        x_t = x0 + t[:, None, None, None] * (spec - x0)
        return self.velocity_fn(x_t, t * self.time_scale_factor, cond), spec - x0, t


class RepetitiveRectifiedFlow(RepeatBinsMixin, RectifiedFlow):
    def __init__(self, vmin: float | int | list, vmax: float | int | list,
                 repeat_bins: int, time_scale_factor=1000,
                 backbone_type=None, backbone_args=None):
        nf, lo, hi = self._geometry(vmin, vmax)
        self.repeat_bins = repeat_bins
        RectifiedFlow.__init__(
            self, out_dims=repeat_bins, num_feats=nf, time_scale_factor=time_scale_factor,
            backbone_type=backbone_type, backbone_args=backbone_args, spec_min=lo, spec_max=hi)


class PitchRectifiedFlow(PitchClipMixin, RepetitiveRectifiedFlow):
    def __init__(self, vmin: float, vmax: float,
                 cmin: float, cmax: float, repeat_bins,
                 time_scale_factor=1000,
                 backbone_type=None, backbone_args=None):
        self.vmin, self.vmax = vmin, vmax
        self.cmin, self.cmax = cmin, cmax
        RepetitiveRectifiedFlow.__init__(
            self, vmin=vmin, vmax=vmax, repeat_bins=repeat_bins, time_scale_factor=time_scale_factor,
            backbone_type=backbone_type, backbone_args=backbone_args)


class MultiVarianceRectifiedFlow(MultiCurveMixin, RepetitiveRectifiedFlow):
    def __init__(
            self, ranges: List[Tuple[float, float]],
            clamps: List[Tuple[float | None, float | None] | None],
            repeat_bins, time_scale_factor=1000,
            backbone_type=None, backbone_args=None
    ):
        assert len(ranges) == len(clamps)
        self.clamps = clamps
        lo, hi = self._ranges(ranges)
        RepetitiveRectifiedFlow.__init__(
            self, vmin=lo, vmax=hi, repeat_bins=repeat_bins, time_scale_factor=time_scale_factor,
            backbone_type=backbone_type, backbone_args=backbone_args)
