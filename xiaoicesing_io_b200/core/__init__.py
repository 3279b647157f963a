from .ddpm import GaussianDiffusion, MultiVarianceDiffusion, PitchDiffusion, RepetitiveDiffusion
from .reflow import (MultiVarianceRectifiedFlow, PitchRectifiedFlow, RectifiedFlow, RepetitiveRectifiedFlow)

__all__ = ['GaussianDiffusion', 'RepetitiveDiffusion', 'PitchDiffusion', 'MultiVarianceDiffusion',
           'RectifiedFlow', 'RepetitiveRectifiedFlow', 'PitchRectifiedFlow', 'MultiVarianceRectifiedFlow']

# the stock inference methods: SamplerBase.forward fuses norm_spec / denorm_spec with the layout changes only around these
from . import _sampling as _s
_s._BASE_INFERENCE.update({GaussianDiffusion.inference, RectifiedFlow.inference})
del _s
