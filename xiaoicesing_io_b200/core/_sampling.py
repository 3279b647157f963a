"""Shared base of GaussianDiffusion / RectifiedFlow on the B200 path: spec (de)normalisation, the
``forward(condition, gt_spec, src_spec, infer)`` entry, buffer + noise plumbing and the program run."""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import nn

from .. import _cabi as C
from ..backbones.wavenet import _time_major_cond
from ..engine import CompiledProgram, allocate_buffers, run_program
from ..schedules import NOISE0, XSTART, Program


def _bounds(values, out_dims):
    """spec_min / spec_max list -> [1, 1, M] (one feature) or [1, F, 1, M] (ddpm.py:103-109)."""
    return torch.FloatTensor(values)[None, None, :out_dims].transpose(-3, -2)


class SamplerBase(nn.Module):
    """Holds the backbone under ``backbone_attr`` ('denoise_fn' / 'velocity_fn') and runs programs."""

    backbone_attr = 'denoise_fn'

    def _init_common(self, out_dims, num_feats, spec_min, spec_max, persistent_bounds):
        self.out_dims = out_dims
        self.num_feats = num_feats
        self.register_buffer('spec_min', _bounds(spec_min, out_dims), persistent=persistent_bounds)
        self.register_buffer('spec_max', _bounds(spec_max, out_dims), persistent=persistent_bounds)
        self._noise_source = None               # test hook: callable(shape) -> N(0,1) tensor [B,F,M,T]

    # ---- spec normalisation (ddpm.py:379-383, reflow.py:140-144) -------------------------------------
    def norm_spec(self, x):
        return (x - self.spec_min) / (self.spec_max - self.spec_min) * 2 - 1

    def denorm_spec(self, x):
        return (x + 1) / 2 * (self.spec_max - self.spec_min) + self.spec_min

    def _source_to_state(self, src_spec):
        """src_spec [B,T,M] / [B,F,T,M] (or curves) -> normalised [B,F,M,T], or None."""
        if src_spec is None:
            return None
        spec = self.norm_spec(src_spec).transpose(-2, -1)
        return spec[:, None] if self.num_feats == 1 else spec

    # ---- program execution -------------------------------------------------------------------------
    def build_program(self) -> Program:                      # pragma: no cover - abstract
        raise NotImplementedError

    def _run(self, cond, b, start, device):
        return sample(getattr(self, self.backbone_attr), self.build_program(), cond, b, self.num_feats,
                      self.out_dims, start, device, noise_source=self._noise_source)

    def _training_forward(self, spec, cond, b, device):     # pragma: no cover - abstract
        raise NotImplementedError

    def forward(self, condition, gt_spec=None, src_spec=None, infer=True):
        """condition [B, T, H] -> de-normalised sample [B, T, M] / [B, F, T, M] (or curves)."""
        cond = condition.transpose(1, 2)                     # [B, H, T] view, as the reference hands it on
        b, device = condition.shape[0], condition.device
        if infer:
            x = self.inference(cond, b, self._source_to_state(src_spec), device)
            return self.denorm_spec(x)
        return self._training_forward(self._source_to_state(gt_spec), cond, b, device)


def sample(backbone, prog: Program, cond_bht: torch.Tensor, b: int, num_feats: int, out_dims: int,
           x_start: Optional[torch.Tensor], device,
           noise_source: Optional[Callable[[tuple], torch.Tensor]] = None) -> torch.Tensor:
    """Runs ``prog`` with ``backbone`` as the denoiser.  Returns [B, T, M] or [B, F, T, M]
    (the transpose of ddpm.py:350 / reflow.py:137 is free: the state is time-major already).

    ``noise_source(shape)`` must return a standard-normal tensor of ``shape`` = (B, F, M, T); the
    default is ``torch.randn`` on ``device``, called in the reference's order (initial draw first,
    then one draw per ancestral step), so a seeded run consumes the same Philox stream as the
    reference on the same device.
    """
    if device is None:
        device = cond_bht.device
    device = torch.device(device)
    if device.type != 'cuda':
        raise C.B2SError(f'sampling needs a CUDA device (got {device}); this path has no CPU fallback')
    B, H, T = cond_bht.shape
    if B != b:
        raise C.B2SError(f'batch size mismatch: cond has {B} utterances, b={b}')
    F_, M = num_feats, out_dims
    shape = (B, F_, M, T)
    if noise_source is None:
        noise_source = lambda s: torch.randn(s, device=device)
    eng = backbone._engine()
    eng.pack()
    cp = CompiledProgram(prog, device)
    sess = eng.begin(_time_major_cond(cond_bht.float()), cp.t_values, per_row_t=False)
    bufs = allocate_buffers(prog, B * T, F_ * M, device)

    def load(src_bfmt, dst):
        C.transpose(src_bfmt.to(device=device, dtype=torch.float32).reshape(B, F_ * M, T).contiguous(), dst,
                    B, F_ * M, T)

    noise0 = noise_source(shape)                        # always drawn first (ddpm.py:227, reflow.py:105)
    if NOISE0 in bufs:
        load(noise0, bufs[NOISE0])
    if prog.needs_x_start:
        assert x_start is not None, 'Missing shallow diffusion source.'
        load(x_start, bufs[XSTART])

    x = run_program(cp, sess, bufs, lambda j, dst: load(noise_source(shape), dst))   # [B*T, F*M]
    x = x.reshape(B, T, F_, M)
    if F_ == 1:
        return x[:, :, 0, :]                            # [B, T, M]
    return x.permute(0, 2, 1, 3).contiguous()           # [B, F, T, M]
