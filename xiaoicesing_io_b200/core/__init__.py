from .ddpm import GaussianDiffusion, MultiVarianceDiffusion, PitchDiffusion, RepetitiveDiffusion
from .reflow import (MultiVarianceRectifiedFlow, PitchRectifiedFlow, RectifiedFlow, RepetitiveRectifiedFlow)

__all__ = ['GaussianDiffusion', 'RepetitiveDiffusion', 'PitchDiffusion', 'MultiVarianceDiffusion',
           'RectifiedFlow', 'RepetitiveRectifiedFlow', 'PitchRectifiedFlow', 'MultiVarianceRectifiedFlow']
