"""Test helpers that build PRODUCT objects (xiaoicesing_io_b200) from golden-fixture metadata."""
from __future__ import annotations

import torch

import xiaoicesing_io_b200 as P


def set_hparams_for(fx):
    m = fx.meta
    P.hparams.clear()
    P.hparams.update(hidden_size=m['hidden_size'], schedule_type='linear', infer=False)
    P.hparams.update(m.get('hparams', {}))


def _ctor_args(ctor):
    c = dict(ctor)
    for k in ('ranges', 'clamps'):
        if k in c:
            c[k] = [None if r is None else tuple(r) for r in c[k]]
    return c


def build_model(fx, device='cpu'):
    """Product sampler module for a 'diffusion' fixture, weights loaded from the fixture."""
    set_hparams_for(fx)
    cls = getattr(P, fx.meta['cls'])
    model = cls(**_ctor_args(fx.meta['ctor']))
    bb = getattr(model, model.backbone_attr)
    bb.load_state_dict(fx.sd, strict=True)
    return model.to(device).eval()


def build_backbone(fx, device='cpu'):
    m = fx.meta
    P.hparams.clear()
    P.hparams.update(hidden_size=m['hidden_size'])
    net = P.build_backbone(m['in_dims'], m['n_feats'], m['backbone_type'], m['backbone_args'])
    net.load_state_dict(fx.sd, strict=True)
    return net.to(device).eval()


def run_program_cpu(prog, denoise, cond, shape, noise0, x_start, step_noise, dtype=torch.float64):
    """Reference executor for a schedules.Program on the CPU (TEST ONLY): buffers in the reference's
    [B,F,M,T] layout, ``lin`` ops as plain tensor arithmetic, ``nfe`` ops through ``denoise``."""
    from xiaoicesing_io_b200.schedules import NOISE0, XSTART
    bufs = {}
    if noise0 is not None:
        bufs[NOISE0] = noise0.to(dtype)
    if x_start is not None:
        bufs[XSTART] = x_start.to(dtype)
    for op in prog.ops:
        if op.kind == 'nfe':
            t = torch.tensor([prog.t_values[op.t_index]], dtype=torch.float32)
            bufs[op.dst] = denoise(bufs[op.src], t, cond).to(dtype)
        elif op.kind == 'lin':
            acc = torch.zeros(shape, dtype=dtype)
            for b, c in op.terms:
                acc = acc + c * bufs[b]
            bufs[op.dst] = acc
        else:
            bufs[op.dst] = step_noise[op.draw].to(dtype)
    return bufs[prog.result]
