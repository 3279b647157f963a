"""Oracle (TEST INFRASTRUCTURE): import the UNMODIFIED reference.

Search order: ``$B2S_REFERENCE_ROOT``, ``/root/reference`` (the build container), ``baseline/_ref`` (the byte-for-byte
copy ``baseline/install_reference.sh`` makes; git-ignored, it travels to the GPU box with the gpurun snapshot).  Used by
``oracle/make_golden.py``, by the optional live cross-check test and by ``bench.py``'s CPU legs (``--impl reference`` and
``cpu_baseline``), which time the reference's own sampling code.  Nothing in the product may import this.

The reference imports ``lightning`` through ``utils/__init__.py:15`` ->
``utils/training_utils.py:7``; that one module is stubbed before import
(SURVEY.md section 8c).  No reference file is modified or copied.
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for cand in (os.environ.get('B2S_REFERENCE_ROOT'), '/root/reference', os.path.join(_HERE, '..', 'baseline', '_ref')):
        if cand and os.path.isfile(os.path.join(cand, 'modules', 'core', 'ddpm.py')):
            return os.path.abspath(cand)
    return os.environ.get('B2S_REFERENCE_ROOT', '/root/reference')


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'modules', 'core', 'ddpm.py'))


_loaded = None


def load():
    """Returns a namespace with the reference's hot-path modules and its global ``hparams`` dict."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')
    if 'utils.training_utils' not in sys.modules:
        stub = types.ModuleType('utils.training_utils')
        stub.get_latest_checkpoint_path = lambda *a, **k: None
        sys.modules['utils.training_utils'] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from utils.hparams import hparams                       # noqa: E402
    hparams.setdefault('hidden_size', 256)
    hparams.setdefault('schedule_type', 'linear')
    hparams.setdefault('infer', False)
    import modules.backbones as backbones                    # noqa: E402
    import modules.core.ddpm as ddpm                         # noqa: E402
    import modules.core.reflow as reflow                     # noqa: E402
    ns = types.SimpleNamespace(hparams=hparams, backbones=backbones, ddpm=ddpm, reflow=reflow)
    _loaded = ns
    return ns


def set_hparams(**kw):
    ref = load()
    ref.hparams.update(kw)
    return ref.hparams


def load_vocoder():
    """The reference's NSF-HiFiGAN generator module (modules/nsf_hifigan/models.py).  Its imports of ``lightning`` (models.py:8,
    only ``rank_zero_info``) and ``matplotlib`` (utils.py:1-3, unused by the generator) are stubbed when those packages are absent;
    no reference file is modified."""
    load()
    import importlib
    for name in ('matplotlib',):
        try:
            importlib.import_module(name)
        except ModuleNotFoundError:
            stub = types.ModuleType(name)
            stub.use = lambda *a, **k: None
            sys.modules[name] = stub
    try:
        importlib.import_module('lightning.pytorch.utilities.rank_zero')
    except ModuleNotFoundError:
        for name in ('lightning', 'lightning.pytorch', 'lightning.pytorch.utilities', 'lightning.pytorch.utilities.rank_zero'):
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules['lightning.pytorch.utilities.rank_zero'].rank_zero_info = print
    import modules.nsf_hifigan.models as models              # noqa: E402
    from modules.nsf_hifigan.env import AttrDict              # noqa: E402
    return types.SimpleNamespace(models=models, AttrDict=AttrDict)


def load_acoustic_infer():
    """The reference's inference driver module (inference/ds_acoustic.py) with its unrelated imports stubbed when absent (``librosa``
    in utils/infer_utils.py:3, ``lightning`` / ``matplotlib`` as for the vocoder); no reference file is modified."""
    load_vocoder()
    import importlib
    class _Anything(types.ModuleType):                          # a stand-in package: any attribute is a callable that must not be called
        __path__ = []

        def __getattr__(self, item):
            if item.startswith('__'):
                raise AttributeError(item)
            return lambda *a, **k: (_ for _ in ()).throw(RuntimeError(f'{self.__name__}.{item} is a stub'))
    for name in ('librosa', 'librosa.filters', 'tqdm', 'onnxruntime', 'scipy.io.wavfile'):
        try:
            importlib.import_module(name)
        except ModuleNotFoundError:
            sys.modules[name] = _Anything(name)
    if isinstance(sys.modules.get('tqdm'), _Anything):
        sys.modules['tqdm'].tqdm = lambda it, **k: it
    import inference.ds_acoustic as ds_acoustic                # noqa: E402
    return ds_acoustic

