import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    # a fresh checkout has no libb2s.so (git-ignored); the package refuses to import without it, so build it first
    # (nvcc cross-compiles without a GPU; a no-op when the library is up to date)
    import __graft_entry__ as entry
    entry._build_module().build()
