"""The acoustic model's inference graph end to end on the B200 kernels: phoneme tokens -> condition (FastSpeech2 acoustic encoder) ->
x_start (ConvNeXt aux decoder, shallow diffusion) -> mel (diffusion / rectified-flow sampling loop) - reference
modules/toplevel.py:32-102 (``DiffSingerAcoustic``), same constructor, same ``hparams`` keys, same sub-module names (``fs2``,
``aux_decoder``, ``diffusion``), so a reference acoustic checkpoint loads with ``load_state_dict(strict=True)``.

``infer=False`` returns the FORWARD values of the reference's training branch (``modules/toplevel.py:103-120``: the aux decoder's
normalised prediction and the denoiser's output on q_sample(gt_mel) with the drawn noise) - enough for validation losses; there is
no autograd through the kernels.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _cabi as C
from .acoustic_encoder import FastSpeech2Acoustic
from .aux_decoder import AuxDecoderAdaptor
from .core import GaussianDiffusion, RectifiedFlow
from .hparams import hparams


class ShallowDiffusionOutput:
    def __init__(self, *, aux_out=None, diff_out=None):
        self.aux_out = aux_out
        self.diff_out = diff_out


def get_backbone_type(root_config: dict, nested_config: dict = None):
    """modules/compat.py:1-11."""
    nested_config = root_config if nested_config is None else nested_config
    return nested_config.get('backbone_type', root_config.get('backbone_type', root_config.get('diff_decoder_type', 'wavenet')))


def get_backbone_args(config: dict, backbone_type: str):
    """modules/compat.py:14-25."""
    args = config.get('backbone_args')
    if args is not None:
        return args
    if backbone_type == 'wavenet':
        return {'num_layers': config.get('residual_layers'), 'num_channels': config.get('residual_channels'),
                'dilation_cycle_length': config.get('dilation_cycle_length')}
    return None


class DiffSingerAcoustic(nn.Module):
    category = 'acoustic'

    def __init__(self, vocab_size, out_dims):
        super().__init__()
        self.fs2 = FastSpeech2Acoustic(vocab_size=vocab_size)
        self.use_shallow_diffusion = hparams.get('use_shallow_diffusion', False)
        self.shallow_args = hparams.get('shallow_diffusion_args', {})
        if self.use_shallow_diffusion:
            self.aux_decoder = AuxDecoderAdaptor(in_dims=hparams['hidden_size'], out_dims=out_dims, num_feats=1,
                                                 spec_min=hparams['spec_min'], spec_max=hparams['spec_max'],
                                                 aux_decoder_arch=self.shallow_args['aux_decoder_arch'],
                                                 aux_decoder_args=self.shallow_args['aux_decoder_args'])
        self.diffusion_type = hparams.get('diffusion_type', 'ddpm')
        self.backbone_type = get_backbone_type(hparams)
        self.backbone_args = get_backbone_args(hparams, self.backbone_type)
        if self.diffusion_type == 'ddpm':
            self.diffusion = GaussianDiffusion(out_dims=out_dims, num_feats=1, timesteps=hparams['timesteps'], k_step=hparams['K_step'],
                                               backbone_type=self.backbone_type, backbone_args=self.backbone_args,
                                               spec_min=hparams['spec_min'], spec_max=hparams['spec_max'])
        elif self.diffusion_type == 'reflow':
            self.diffusion = RectifiedFlow(out_dims=out_dims, num_feats=1, t_start=hparams['T_start'],
                                           time_scale_factor=hparams['time_scale_factor'], backbone_type=self.backbone_type,
                                           backbone_args=self.backbone_args, spec_min=hparams['spec_min'], spec_max=hparams['spec_max'])
        else:
            raise NotImplementedError(self.diffusion_type)

    @torch.no_grad()
    def forward(self, txt_tokens, mel2ph, f0, key_shift=None, speed=None, spk_embed_id=None, gt_mel=None, infer=True, **kwargs
                ) -> ShallowDiffusionOutput:
        condition = self.fs2(txt_tokens, mel2ph, f0, key_shift=key_shift, speed=speed, spk_embed_id=spk_embed_id, **kwargs)
        if not infer:                                                      # forward values only: validation losses
            if gt_mel is None:
                raise C.B2SError('infer=False needs gt_mel (modules/toplevel.py:103-120)')
            aux_out = diff_out = None
            if self.use_shallow_diffusion:
                if self.shallow_args.get('train_aux_decoder', True):
                    g = float(self.shallow_args.get('aux_decoder_grad', 1.0))
                    aux_out = self.aux_decoder(condition * g + condition * (1 - g), infer=False)          # :108 (values; no autograd)
                if self.shallow_args.get('train_diffusion', True):
                    diff_out = self.diffusion(condition, gt_spec=gt_mel, infer=False)
            else:
                diff_out = self.diffusion(condition, gt_spec=gt_mel, infer=False)
            return ShallowDiffusionOutput(aux_out=aux_out, diff_out=diff_out)
        keep = (mel2ph > 0).float()[:, :, None]                            # modules/toplevel.py:95, :101
        if self.use_shallow_diffusion:
            aux_mel_pred = self.aux_decoder(condition, infer=True)
            aux_mel_pred *= keep
            src_mel = gt_mel if (gt_mel is not None and self.shallow_args.get('val_gt_start')) else aux_mel_pred
        else:
            aux_mel_pred = src_mel = None
        mel_pred = self.diffusion(condition, src_spec=src_mel, infer=True)
        mel_pred *= keep
        return ShallowDiffusionOutput(aux_out=aux_mel_pred, diff_out=mel_pred)
