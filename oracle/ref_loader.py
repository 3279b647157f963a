"""Oracle (TEST INFRASTRUCTURE): import the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
``oracle/make_golden.py`` and by the optional live cross-check test.  Nothing in
the product, ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.

The reference imports ``lightning`` through ``utils/__init__.py:15`` ->
``utils/training_utils.py:7``; that one module is stubbed before import
(SURVEY.md section 8c).  No reference file is modified or copied.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get('B2S_REFERENCE_ROOT', '/root/reference')


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
