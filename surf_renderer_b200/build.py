"""Build libsurf_b200.so (sm_100a) in-tree with nvcc.  `python -m surf_renderer_b200.build [--force] [-v]`.

The library is five translation units (csrc/surf_launch.cuh lists them) compiled in parallel to objects under
csrc/_obj/ and linked into one shared library; only the units whose sources changed are recompiled."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(CSRC, '_obj')
SO = os.path.join(PKG, 'libsurf_b200.so')
HEADER = os.path.join(PKG, '..', 'include', 'surf_b200.h')
COMMON = ['surf_view.h', 'surf_math.cuh', 'surf_runtime.cuh', 'surf_ptx.cuh', 'surf_batch.cuh', 'surf_launch.cuh']
ISECT = COMMON + ['surf_intersect.cuh']
# translation unit -> headers it includes (csrc-relative)
UNITS = {
    'surf_kernels.cu': COMMON + ['surf_frame_kernels.cuh', 'surf_shade.cuh', 'surf_backward.cuh', 'surf_splats.cuh',
                                 'surf_fast.cuh', 'surf_scatter.cuh'],
    'surf_isect_main.cu': ISECT,
    'surf_isect_batch.cu': ISECT,
    'surf_isect_const.cu': ISECT,
    'surf_isect_rays.cu': ISECT + ['surf_intersect_rays.cuh'],
}
SOURCES = [os.path.join(CSRC, u) for u in UNITS]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '--expt-relaxed-constexpr', '-fmad=false', '-Xcompiler', '-fPIC,-ffp-contract=off']


def nvcc_path():
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: libsurf_b200.so cannot be built')


def _deps(unit):
    return [os.path.join(CSRC, unit), HEADER, os.path.abspath(__file__)] + \
        [os.path.join(CSRC, h) for h in UNITS[unit] if os.path.exists(os.path.join(CSRC, h))]


def _obj(unit):
    return os.path.join(OBJ, unit.replace('.cu', '.o'))


def _unit_fresh(unit):
    o = _obj(unit)
    return os.path.exists(o) and all(os.path.getmtime(o) >= os.path.getmtime(d) for d in _deps(unit))


def is_fresh():
    return os.path.exists(SO) and all(_unit_fresh(u) and os.path.getmtime(SO) >= os.path.getmtime(_obj(u)) for u in UNITS)


def build_variant(out_so, defines, units=('surf_kernels.cu',)):
    """A/B build: recompile `units` with extra -D defines into a separate object directory and link `out_so` from
    them plus the regular objects of the other units (load it with SURF_B200_LIB=<out_so>)."""
    build()
    nvcc = nvcc_path()
    vdir = out_so + '.obj'
    os.makedirs(vdir, exist_ok=True)
    objs = []
    for u in UNITS:
        if u in units:
            o = os.path.join(vdir, u.replace('.cu', '.o'))
            subprocess.check_call([nvcc] + NVCC_FLAGS + ['-D' + d for d in defines] + ['-c', os.path.join(CSRC, u), '-o', o])
            objs.append(o)
        else:
            objs.append(_obj(u))
    subprocess.check_call([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a'] + objs + ['-o', out_so])
    return out_so


def build(force=False, verbose=False):
    if not force and is_fresh():
        return SO
    os.makedirs(OBJ, exist_ok=True)
    nvcc = nvcc_path()
    todo = [u for u in UNITS if force or not _unit_fresh(u)]

    def compile_unit(unit):
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, unit), '-o', _obj(unit)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return unit, r

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        results = list(ex.map(compile_unit, todo))
    for unit, r in results:
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError('nvcc failed on %s' % unit)
    subprocess.check_call([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a'] + [_obj(u) for u in UNITS] + ['-o', SO])
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
