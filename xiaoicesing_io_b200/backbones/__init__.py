"""Backbone registry + factory: the plug-in point (reference modules/backbones/__init__.py:6-18)."""
import inspect

import torch.nn

from .lynxnet import LYNXNet
from .wavenet import WaveNet

BACKBONES = {
    'wavenet': WaveNet,
    'lynxnet': LYNXNet
}


def filter_kwargs(dict_to_filter, kwarg_obj):
    """Drops kwargs the constructor does not accept, e.g. the YAML's ``dropout_rate``
    (reference utils/__init__.py:149-163)."""
    sig = inspect.signature(kwarg_obj)
    if any(param.kind == param.VAR_KEYWORD for param in sig.parameters.values()):
        return dict_to_filter.copy()
    keys = [p.name for p in sig.parameters.values()
            if p.kind == p.POSITIONAL_OR_KEYWORD or p.kind == p.KEYWORD_ONLY]
    return {k: dict_to_filter[k] for k in keys if k in dict_to_filter}


def build_backbone(out_dims: int, num_feats: int, backbone_type: str, backbone_args: dict) -> torch.nn.Module:
    backbone = BACKBONES[backbone_type]
    kwargs = filter_kwargs(backbone_args, backbone)
    return BACKBONES[backbone_type](out_dims, num_feats, **kwargs)
