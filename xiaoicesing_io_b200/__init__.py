"""xiaoicesing_io_b200 - B200-native (sm_100a) implementation of ONE hot path of the DiffSinger fork
``vsingerxiaoice-rwkv/xiaoicesing-io``: the repeated denoiser forward inside the diffusion /
rectified-flow sampling loop over mel tensors [B, 128 bins, T frames].

Public surface (same names and contracts as the reference's ``modules.backbones`` / ``modules.core``):

    BACKBONES, build_backbone, WaveNet, LYNXNet
    GaussianDiffusion, RepetitiveDiffusion, PitchDiffusion, MultiVarianceDiffusion
    RectifiedFlow, RepetitiveRectifiedFlow, PitchRectifiedFlow, MultiVarianceRectifiedFlow
    AUX_DECODERS, build_aux_decoder, ConvNeXtDecoder, AuxDecoderAdaptor   (modules.aux_decoder: the producer of x_start)
    FastSpeech2Acoustic, FastSpeech2Encoder   (modules.fastspeech: the producer of the condition tensor, rotary configuration)
    DiffSingerAcoustic, ShallowDiffusionOutput   (modules.toplevel: tokens -> condition -> x_start -> mel, inference)
    vocoder.Generator, vocoder.load_model, NsfHifiGAN   (modules.nsf_hifigan / modules.vocoders: mel + f0 -> waveform)
    DiffusionLoss, RectifiedFlowLoss   (modules.losses: forward values of the training-branch losses, validation)
    infer.DiffSingerAcousticInfer   (inference.ds_acoustic: .ds segment -> model inputs -> mel -> waveform / .mel.pt)
    hparams  (the global config dict, utils/hparams.py:13)
    segments (batched .ds segment driver: ragged batches, per-segment seeds, .mel.pt writer), partition (multi-GPU), B2SError

Importing the package loads ``libb2s.so`` (hand-written CUDA kernels behind the C ABI of
``include/b2s.h``); it raises ImportError if the library has not been built.  There is no CPU or
PyTorch fallback anywhere in this package.
"""
from . import _cabi  # noqa: F401  (fails loudly when libb2s.so is missing)
from ._cabi import B2SError
from .backbones import BACKBONES, LYNXNet, WaveNet, build_backbone, filter_kwargs
from .core import (GaussianDiffusion, MultiVarianceDiffusion, MultiVarianceRectifiedFlow, PitchDiffusion,
                   PitchRectifiedFlow, RectifiedFlow, RepetitiveDiffusion, RepetitiveRectifiedFlow)
from .hparams import hparams, set_hparams
from .aux_decoder import AUX_DECODERS, AuxDecoderAdaptor, ConvNeXtDecoder, build_aux_decoder
from .acoustic_encoder import FastSpeech2Acoustic, FastSpeech2Encoder
from .toplevel import DiffSingerAcoustic, ShallowDiffusionOutput
from . import infer, losses, partition, segments, vocoder  # noqa: E402,F401
from .losses import DiffusionLoss, RectifiedFlowLoss
from .vocoder import NsfHifiGAN

__version__ = '0.2.0'
