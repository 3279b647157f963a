"""In-tree build of libb2s.so (hand-written sm_100a CUDA + the C ABI of include/b2s.h).

``nvcc`` cross-compiles without a GPU; the resulting ``.so`` lives next to this file (git-ignored,
but it travels with the gpurun snapshot).  No torch headers are involved: the boundary is plain C.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, os.environ.get('B2S_LIB_OUT', 'libb2s.so'))      # B2S_LIB_OUT: side builds (profiling variants)
OBJ_DIR = os.path.join(HERE, 'csrc', '_obj' + ('_tlog' if os.environ.get('B2S_BUILD_TLOG') else '') + ('_exp' if os.environ.get('B2S_BUILD_EXPERIMENTS') else ''))

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-Wall',
    '--expt-relaxed-constexpr',
]


if os.environ.get('B2S_BUILD_EXPERIMENTS'):   # measured-and-rejected kernel variants: transposed stack kernel, single-CTA
    NVCC_FLAGS.append('-DB2S_EXPERIMENTS')    # fused layer / GEMM, cta_group::2 variant of the round-1 stack kernel, update inside the denoiser launch
if os.environ.get('B2S_BUILD_TLOG'):          # profiling build: in-kernel phase timestamps (scripts/stack_timeline.py)
    NVCC_FLAGS.append('-DB2S_TLOG')


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; libb2s.so cannot be built')


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles every ``csrc/*.cu`` for sm_100a and links ``libb2s.so``.  Returns the library path."""
    srcs = sources()
    headers = glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(HERE, '..', 'include', '*.h'))
    if not force and not _stale(LIB_PATH, srcs + headers):
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *NVCC_FLAGS, '-Xptxas', '-v', '-c', src, '-o', obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
            with open(obj + '.ptxas.log', 'w') as f:
                f.write(r.stderr)
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, '-shared', '-o', LIB_PATH, *objs, '-lcudart']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB_PATH


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose=True))
