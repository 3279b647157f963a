"""Repeat-bin wrappers shared by the diffusion and rectified-flow samplers.

The variance models predict scalar curves (pitch delta, energy, breathiness, ...).  A curve is tiled
over ``repeat_bins`` pseudo-mel bins before normalisation, sampled like a mel, and averaged over the
bins afterwards; pitch is clipped to [cmin, cmax], multi-variance outputs are clamped per curve.
The reference states this twice (ddpm.py:386-505 and reflow.py:147-261); here it is one set of mixins
applied to both sampler classes.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def _scalar(v) -> bool:
    return isinstance(v, (int, float))


class RepeatBinsMixin:
    """``x [B,T]`` or ``[B,F,T]``  <->  ``[B,T,R]`` or ``[B,F,T,R]`` around the base norm/denorm."""

    repeat_bins: int

    @staticmethod
    def _geometry(vmin, vmax):
        assert (_scalar(vmin) and _scalar(vmax)) or len(vmin) == len(vmax)
        nf = 1 if _scalar(vmin) else len(vmin)
        lo = [vmin] if nf == 1 else [[v] for v in vmin]
        hi = [vmax] if nf == 1 else [[v] for v in vmax]
        return nf, lo, hi

    def norm_spec(self, x):
        tiled = x.unsqueeze(-1).expand(*x.shape, self.repeat_bins)
        return super().norm_spec(tiled)

    def denorm_spec(self, x):
        return super().denorm_spec(x).mean(dim=-1)


class PitchClipMixin:
    """Clip the curve before normalising and after de-normalising (ddpm.py:441-445)."""
    cmin: float
    cmax: float

    def norm_spec(self, x):
        return super().norm_spec(x.clamp(min=self.cmin, max=self.cmax))

    def denorm_spec(self, x):
        return super().denorm_spec(x).clamp(min=self.cmin, max=self.cmax)


class MultiCurveMixin:
    """A list of F curves, each with an optional (lo, hi) clamp (ddpm.py:471-505)."""
    clamps: Sequence[Optional[Tuple[Optional[float], Optional[float]]]]

    @staticmethod
    def _ranges(ranges):
        lo = [r[0] for r in ranges]
        hi = [r[1] for r in ranges]
        if len(lo) == 1:
            return lo[0], hi[0]
        return lo, hi

    def clamp_spec(self, xs) -> List[torch.Tensor]:
        return [x if c is None else x.clamp(min=c[0], max=c[1]) for x, c in zip(xs, self.clamps)]

    def norm_spec(self, xs):
        assert len(xs) == self.num_feats
        stacked = torch.stack(self.clamp_spec(xs), dim=1)            # [B, F, T]
        if self.num_feats == 1:
            stacked = stacked.squeeze(1)
        return super().norm_spec(stacked)

    def denorm_spec(self, xs):
        curves = super().denorm_spec(xs)                              # [B, T] or [B, F, T]
        curves = [curves] if self.num_feats == 1 else list(curves.unbind(dim=1))
        assert len(curves) == self.num_feats
        return self.clamp_spec(curves)
