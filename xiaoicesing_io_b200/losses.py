"""Training-branch losses, FORWARD VALUES ONLY (validation during training, SURVEY.md section 3.3 / section 8 row f-4's forward half;
reference modules/losses/diff_loss.py:5-37, modules/losses/reflow_loss.py:6-50): same class names, constructor arguments and call
signatures.  One masked-reduction launch (b2s_masked_loss_f32) instead of the reference's five elementwise passes; deterministic.
There is no backward: these modules return a detached scalar, and the package has no autograd path."""
from __future__ import annotations

import torch
from torch import nn

from . import _cabi as C


def _loss(a, b, non_padding, t, l1):
    a = C.require_cuda(a.detach().float().contiguous(), 'prediction')
    b = C.require_cuda(b.detach().float().contiguous(), 'target')
    if a.shape != b.shape or a.dim() != 4:
        raise C.B2SError(f'prediction {tuple(a.shape)} / target {tuple(b.shape)}: expected two [B, F, M, T] tensors')
    B, F, M, T = a.shape
    if non_padding is not None:
        non_padding = C.require_cuda(non_padding.detach().float().contiguous(), 'non_padding')
        if non_padding.dim() != 3 or non_padding.shape[0] != B or non_padding.shape[1] != T or non_padding.shape[2] not in (1, M):
            raise C.B2SError(f'non_padding {tuple(non_padding.shape)}: expected [B, T, 1] or [B, T, M]')
    if t is not None:
        t = C.require_cuda(t.detach().float().contiguous(), 't')
        if t.shape != (B,):
            raise C.B2SError(f't {tuple(t.shape)}: expected [B]')
    out = torch.empty(1, device=a.device)
    with torch.cuda.device(a.device):
        C.masked_loss(a, b, non_padding, t, l1, out)
    return out[0]


class DiffusionLoss(nn.Module):
    """modules/losses/diff_loss.py:5-37: ``forward(x_recon [B, 1, M, T], noise, non_padding [B, T, M] = None) -> scalar``."""

    def __init__(self, loss_type):
        super().__init__()
        if loss_type not in ('l1', 'l2'):
            raise NotImplementedError()
        self.loss_type = loss_type

    def forward(self, x_recon, noise, non_padding=None):
        return _loss(x_recon, noise, non_padding, None, self.loss_type == 'l1')


class RectifiedFlowLoss(nn.Module):
    """modules/losses/reflow_loss.py:6-50: ``forward(v_pred, v_gt, t [B], non_padding = None) -> scalar``, log-normal time weights."""

    def __init__(self, loss_type, log_norm=True):
        super().__init__()
        if loss_type not in ('l1', 'l2'):
            raise NotImplementedError()
        self.loss_type, self.log_norm = loss_type, log_norm

    def forward(self, v_pred, v_gt, t, non_padding=None):
        return _loss(v_pred, v_gt, non_padding, t if self.log_norm else None, self.loss_type == 'l1')
