"""Build the CPU emulation of the kernel math (test infrastructure, see surf_emul.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'libsurf_emul.so')
SRC = os.path.join(HERE, 'surf_emul.cpp')
DEPS = [SRC] + [os.path.join(HERE, '..', '..', 'surf_renderer_b200', 'csrc', f) for f in ('surf_math.cuh', 'surf_view.h')] \
    + [os.path.join(HERE, '..', '..', 'include', 'surf_b200.h')]


def build(force=False):
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in DEPS):
        return SO
    cmd = ['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC', '-x', 'c++', SRC, '-o', SO]
    subprocess.check_call(cmd)
    return SO


if __name__ == '__main__':
    print(build(force=True))
