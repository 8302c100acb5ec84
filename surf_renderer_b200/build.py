"""Build libsurf_b200.so (sm_100a) in-tree with nvcc.  `python -m surf_renderer_b200.build [--force]`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
SO = os.path.join(PKG, 'libsurf_b200.so')
SOURCES = [os.path.join(CSRC, 'surf_kernels.cu')]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))] + \
    [os.path.join(PKG, '..', 'include', 'surf_b200.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '--expt-relaxed-constexpr', '-fmad=false', '-shared', '-Xcompiler', '-fPIC,-ffp-contract=off']


def nvcc_path():
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: libsurf_b200.so cannot be built')


def is_fresh():
    return os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in DEPS)


def build(force=False, verbose=False):
    if not force and is_fresh():
        return SO
    cmd = [nvcc_path()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + SOURCES + ['-o', SO]
    subprocess.check_call(cmd)
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
