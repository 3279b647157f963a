"""Helpers shared by the oracle tests and the GPU parity tests: fixture loading and
running the ORACLE on a fixture (test infrastructure only)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np
import torch

from oracle import denoisers as OD
from oracle import samplers as OS

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def fixture_names(prefix=''):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, f'{prefix}*.npz')))
    return [n for n in names if not n.startswith('weights_')]


class Fixture:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.name = name
        self.meta = json.loads(bytes(z['meta']).decode())
        self.arrays = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('in.')}
        w = np.load(os.path.join(GOLDEN_DIR, self.meta['weights'] + '.npz'))
        self.sd = {k: torch.from_numpy(w[k]) for k in w.files}

    def __getitem__(self, k):
        return self.arrays[k]

    def __contains__(self, k):
        return k in self.arrays


def backbone_cfg(backbone_type, backbone_args, in_dims, n_feats, hidden_size):
    if backbone_type == 'wavenet':
        keys = ('num_layers', 'num_channels', 'dilation_cycle_length')
        return OD.WaveNetCfg(in_dims=in_dims, n_feats=n_feats, hidden_size=hidden_size,
                             **{k: backbone_args[k] for k in keys if k in backbone_args})
    keys = ('num_layers', 'num_channels', 'expansion_factor', 'kernel_size', 'activation', 'strong_cond')
    return OD.LYNXNetCfg(in_dims=in_dims, n_feats=n_feats, hidden_size=hidden_size,
                         **{k: backbone_args[k] for k in keys if k in backbone_args})


def variance_geometry(meta):
    """(out_dims, num_feats, spec_min, spec_max, clamp_in, clamp_out) for the class of a diffusion fixture,
    following ddpm.py:386-505 / reflow.py:147-261."""
    c = meta['ctor']
    cls = meta['cls']
    if cls in ('GaussianDiffusion', 'RectifiedFlow'):
        smin = torch.tensor(c['spec_min'], dtype=torch.float32)[None, None, :c['out_dims']].transpose(-3, -2)
        smax = torch.tensor(c['spec_max'], dtype=torch.float32)[None, None, :c['out_dims']].transpose(-3, -2)
        return c['out_dims'], c.get('num_feats', 1), smin, smax, None
    if cls.startswith('Pitch'):
        vmin, vmax, clamps = c['vmin'], c['vmax'], [(c['cmin'], c['cmax'])]
    else:
        vmin = [r[0] for r in c['ranges']]
        vmax = [r[1] for r in c['ranges']]
        clamps = c['clamps']
        if len(vmin) == 1:
            vmin, vmax = vmin[0], vmax[0]
    nf = 1 if isinstance(vmin, (int, float)) else len(vmin)
    smin = [vmin] if nf == 1 else [[v] for v in vmin]
    smax = [vmax] if nf == 1 else [[v] for v in vmax]
    R = c['repeat_bins']
    smin = torch.tensor(smin, dtype=torch.float32)[None, None, :R].transpose(-3, -2)
    smax = torch.tensor(smax, dtype=torch.float32)[None, None, :R].transpose(-3, -2)
    return R, nf, smin, smax, clamps


def run_oracle_backbone(fx: Fixture, dtype=torch.float32):
    m = fx.meta
    cfg = backbone_cfg(m['backbone_type'], m['backbone_args'], m['in_dims'], m['n_feats'], m['hidden_size'])
    fn = OD.make_denoiser(fx.sd, cfg, dtype)
    return fn(fx['spec'], fx['t'], fx['cond'])


def run_oracle_diffusion(fx: Fixture, dtype=torch.float32, denoise=None):
    """Restates ``model(condition, src_spec, infer=True)`` for every class in the fixtures, with the
    fixture's recorded random draws.  ``denoise`` overrides the backbone (used to drive the PRODUCT's
    sampler logic tests with the oracle denoiser is done elsewhere; here it is the oracle's own)."""
    m = fx.meta
    c, hp = m['ctor'], m['hparams']
    M, nf, smin, smax, clamps = variance_geometry(m)
    cfg = backbone_cfg(c['backbone_type'], c['backbone_args'], M, nf, m['hidden_size'])
    if denoise is None:
        denoise = OD.make_denoiser(fx.sd, cfg, dtype)
    cond = fx['condition'].transpose(1, 2).to(dtype)
    draws = fx['draws']
    is_reflow = 'Rectified' in m['cls']
    # ---- src_spec -> normalised x_start [B,F,M,T]
    src = fx['src_spec'] if 'src_spec' in fx else None
    x_start = None
    if src is not None:
        spec = OS.norm_spec(src.to(dtype), smin.to(dtype), smax.to(dtype)).transpose(-2, -1)
        x_start = spec[:, None] if nf == 1 else spec
    if is_reflow:
        use_shallow = hp.get('use_shallow_diffusion', False)
        t_start = c.get('t_start', 0.) if use_shallow else 0.
        x = OS.rectified_flow_inference(
            denoise, cond, t_start=hp.get('T_start_infer', t_start), use_shallow=use_shallow,
            algorithm=hp['sampling_algorithm'], steps=hp['sampling_steps'], noise0=draws[0],
            time_scale_factor=c.get('time_scale_factor', 1000), x_end=x_start, dtype=dtype)
    else:
        use_shallow = hp.get('use_shallow_diffusion', False)
        timesteps = c.get('timesteps', 1000)
        k_step = c.get('k_step', 1000) if use_shallow else timesteps
        sch = OS.DiffusionSchedule(timesteps, hp.get('schedule_type', 'linear'), dtype=dtype)
        x = OS.gaussian_diffusion_inference(
            denoise, sch, cond, k_step=k_step, timesteps=timesteps, use_shallow=use_shallow,
            K_step_infer=hp.get('K_step_infer', k_step), speedup=hp['diff_speedup'],
            accelerator=hp['diff_accelerator'], noise0=draws[0], x_start=x_start,
            step_noise=draws[1:], dtype=dtype)
    out = OS.denorm_spec(x, smin.to(dtype), smax.to(dtype))
    if clamps is None:
        return out
    out = out.mean(dim=-1)                      # RepetitiveDiffusion.denorm_spec, ddpm.py:415-421
    outs = [out] if nf == 1 else list(out.unbind(dim=1))
    res = []
    for o, cl in zip(outs, clamps):
        res.append(o if cl is None else o.clamp(min=cl[0], max=cl[1]))
    if m['cls'].startswith('Pitch'):
        return res[0]
    return res


def expected_outputs(fx: Fixture):
    if 'out' in fx:
        return fx['out']
    outs = []
    i = 0
    while f'out{i}' in fx:
        outs.append(fx[f'out{i}'])
        i += 1
    return outs
