"""GaussianDiffusion and its repeat-bin variants on the B200 path.

Drop-in for the reference's modules/core/ddpm.py:55-505: same class names, constructor signatures,
registered buffer names (so checkpoints load), ``forward`` / ``inference`` contracts, hparams keys
(K_step_infer, diff_speedup, diff_accelerator, schedule_type, use_shallow_diffusion) and error
behaviour.  The K-step loop itself is compiled by ``schedules`` and run by libb2s kernels.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from .. import schedules as S
from ..backbones import build_backbone
from ..hparams import hparams
from ._sampling import SamplerBase
from ._variants import MultiCurveMixin, PitchClipMixin, RepeatBinsMixin


class GaussianDiffusion(SamplerBase):
    backbone_attr = 'denoise_fn'

    def __init__(self, out_dims, num_feats=1, timesteps=1000, k_step=1000,
                 backbone_type=None, backbone_args=None, betas=None,
                 spec_min=None, spec_max=None):
        super().__init__()
        self.denoise_fn = build_backbone(out_dims, num_feats, backbone_type, backbone_args)
        if betas is None:
            betas = S.beta_schedule[hparams['schedule_type']](timesteps)
        elif isinstance(betas, torch.Tensor):
            betas = betas.detach().cpu().numpy()
        self._tables = S.DiffusionTables(betas)
        self.use_shallow_diffusion = hparams.get('use_shallow_diffusion', False)
        if self.use_shallow_diffusion:
            assert k_step <= timesteps, 'K_step should not be larger than timesteps.'
        self.timesteps = timesteps
        self.k_step = k_step if self.use_shallow_diffusion else timesteps
        for name, buf in self._tables.fp32().items():          # same names + persistence as ddpm.py:82-101
            self.register_buffer(name, buf)
        self._init_common(out_dims, num_feats, spec_min, spec_max, persistent_bounds=True)
        # attributes the ONNX exporter reads (ddpm.py:111-115)
        self.time_scale_factor = self.timesteps
        self.t_start = 1 - self.k_step / self.timesteps
        divisors = [i for i in range(1, self.timesteps + 1) if self.timesteps % i == 0]
        self.register_buffer('timestep_factors', torch.LongTensor(divisors), persistent=False)

    def _tables_current(self) -> S.DiffusionTables:
        """Host tables re-derived from the ``betas`` buffer if a checkpoint replaced it."""
        b32 = self.betas.detach().cpu().numpy()
        if not np.array_equal(b32, self._tables.betas.astype(np.float32)):
            self._tables = S.DiffusionTables(b32.astype(np.float64))
        return self._tables

    def build_program(self) -> S.Program:
        return S.build_gaussian_program(
            self._tables_current(), self.betas.detach().cpu(), timesteps=self.timesteps, k_step=self.k_step,
            use_shallow=self.use_shallow_diffusion, K_step_infer=hparams.get('K_step_infer', self.k_step),
            speedup=hparams['diff_speedup'], accelerator=hparams.get('diff_accelerator'))

    @torch.no_grad()
    def inference(self, cond, b=1, x_start=None, device=None, lengths=None, initial_noise=None):
        """cond [B, H, T]; x_start normalised [B, F, M, T] or None  ->  [B, T, M] / [B, F, T, M]."""
        return self._run(cond, b, x_start, device, lengths, initial_noise)

    @torch.no_grad()
    def _training_forward(self, spec, cond, b, device):
        """``forward(infer=False)``: one denoiser call on q_sample(x0, t, noise) with t ~ U{0..k_step-1}
        (ddpm.py:206-219, 360-367).  Runs the inference kernels, so it serves validation losses; gradient
        training of the denoiser is outside this path."""
        t = torch.randint(0, self.k_step, (b,), device=device).long()
        noise = torch.randn_like(spec)
        view = (b,) + (1,) * (spec.dim() - 1)
        x_noisy = (self.sqrt_alphas_cumprod[t].reshape(view) * spec
                   + self.sqrt_one_minus_alphas_cumprod[t].reshape(view) * noise)
        return self.denoise_fn(x_noisy, t, cond), noise


class RepetitiveDiffusion(RepeatBinsMixin, GaussianDiffusion):
    def __init__(self, vmin: float | int | list, vmax: float | int | list,
                 repeat_bins: int, timesteps=1000, k_step=1000,
                 backbone_type=None, backbone_args=None,
                 betas=None):
        nf, lo, hi = self._geometry(vmin, vmax)
        self.repeat_bins = repeat_bins
        GaussianDiffusion.__init__(
            self, out_dims=repeat_bins, num_feats=nf, timesteps=timesteps, k_step=k_step,
            backbone_type=backbone_type, backbone_args=backbone_args, betas=betas, spec_min=lo, spec_max=hi)


class PitchDiffusion(PitchClipMixin, RepetitiveDiffusion):
    def __init__(self, vmin: float, vmax: float,
                 cmin: float, cmax: float, repeat_bins,
                 timesteps=1000, k_step=1000,
                 backbone_type=None, backbone_args=None,
                 betas=None):
        self.vmin, self.vmax = vmin, vmax          # normalisation range
        self.cmin, self.cmax = cmin, cmax          # clipping range
        RepetitiveDiffusion.__init__(
            self, vmin=vmin, vmax=vmax, repeat_bins=repeat_bins, timesteps=timesteps, k_step=k_step,
            backbone_type=backbone_type, backbone_args=backbone_args, betas=betas)


class MultiVarianceDiffusion(MultiCurveMixin, RepetitiveDiffusion):
    def __init__(
            self, ranges: List[Tuple[float, float]],
            clamps: List[Tuple[float | None, float | None] | None],
            repeat_bins, timesteps=1000, k_step=1000,
            backbone_type=None, backbone_args=None,
            betas=None
    ):
        assert len(ranges) == len(clamps)
        self.clamps = clamps
        lo, hi = self._ranges(ranges)
        RepetitiveDiffusion.__init__(
            self, vmin=lo, vmax=hi, repeat_bins=repeat_bins, timesteps=timesteps, k_step=k_step,
            backbone_type=backbone_type, backbone_args=backbone_args, betas=betas)
