"""In-tree build of ``libb200remap.so`` (sm_100a only) with nvcc.

The shared object is written next to the package (``pyremap_b200/_lib``) so it
travels with the tree; it is git-ignored.  nvcc cross-compiles without a GPU.
"""

from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.path.join(_PKG, 'csrc', 'b200remap.cu')
HEADER = os.path.join(os.path.dirname(_PKG), 'include', 'b200remap.h')
LIB_DIR = os.path.join(_PKG, '_lib')
LIB_PATH = os.path.join(LIB_DIR, 'libb200remap.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17', '-fmad=false',
    '--shared', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'),
                 '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: libb200remap.so cannot be built')


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(f) and os.path.getmtime(f) > built
               for f in (SOURCE, HEADER))


def build_library(force=False, verbose=False, extra_flags=()):
    """Compile the CUDA library if missing or older than its sources."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = LIB_PATH + '.tmp'
    cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, '-o', tmp, SOURCE]
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == '__main__':
    import sys
    print(build_library(force='--force' in sys.argv, verbose=True,
                        extra_flags=['-Xptxas', '-v'] if '-v' in sys.argv else ()))
