"""CPU test: libb2s.so loads, exports every symbol include/b2s.h declares, and its argument
validation answers without touching a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(experiments=False):
    text = open(os.path.join(ROOT, 'include', 'b2s.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    if not experiments:                      # declarations of B2S_BUILD_EXPERIMENTS=1 builds only
        text = re.sub(r'#ifdef B2S_EXPERIMENTS.*?#endif', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(b2s_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    from xiaoicesing_io_b200 import _cabi
    return _cabi


def test_every_declared_symbol_is_exported(lib):
    names = _declared(experiments=lib.HAS_EXPERIMENTS)
    assert len(names) >= 10
    raw = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f'{n} declared in include/b2s.h but not exported by libb2s.so'
    table = set(lib.EXPORTED_SYMBOLS) | (set(lib.EXPERIMENTAL) if lib.HAS_EXPERIMENTS else set())
    assert set(names) == table, ('ctypes table and header disagree', set(names) ^ table)
    assert set(_declared(True)) - set(_declared(False)) == set(lib.EXPERIMENTAL)


def test_abi_version(lib):
    assert lib.lib.b2s_abi_version() == 5


def test_argument_validation_needs_no_gpu(lib):
    rc = lib.lib.b2s_sampler_lincomb_f32(None, None, None, 1, 16, None)
    assert rc == -1
    assert b'null pointer' in lib.lib.b2s_last_error()
    rc = lib.lib.b2s_wavenet_gate_f32(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 4,
                                      ctypes.c_void_p(16), 1, 8, 20, 1, None)
    assert rc == -1 and b'multiple of 16' in lib.lib.b2s_last_error()
    with pytest.raises(lib.B2SError):
        lib.check(rc, 'b2s_wavenet_gate_f32')
