"""Copy recipe for oracle/_ref/: the UNMODIFIED reference files of the render path, so that the checker and
`bench.py --impl reference` can run the reference's own `diffrend.torch.renderer.render` where /root/reference does
not exist (the GPU box).  Test infrastructure only: nothing under surf_renderer_b200/ imports it.

    python oracle/make_ref.py [--src /root/reference]

oracle/_ref/ is git-ignored (the reference's sources never enter this repository's history) but not gpurun-ignored,
so it travels to the GPU box like the built .so.  The reference is pure Python: there is nothing to compile.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
# files of the render path and of the loaders its scenes come from; sub-directories without __init__.py
# are implicit namespace packages in the reference, and stay so here
FILES = [
    'diffrend/__init__.py',
    'diffrend/torch/renderer.py', 'diffrend/torch/utils.py', 'diffrend/torch/ops.py', 'diffrend/torch/params.py',
    'diffrend/torch/render.py', 'diffrend/torch/projection_layer.py',
    'diffrend/utils/utils.py', 'diffrend/utils/sample_generator.py',
    'diffrend/model.py',
    'diffrend/numpy/renderer.py', 'diffrend/numpy/ops.py', 'diffrend/numpy/quaternion.py', 'diffrend/numpy/vector.py',
    'data/__init__.py', 'data/bunny.splat', 'data/torus_1K.obj', 'data/chair_0001.off', 'data/cube.obj',
    'scenes/basic.json', 'scenes/halfbox_sphere_cube.json',
    'scenes/objs/triangle2d.obj', 'scenes/objs/halfbox.obj', 'scenes/objs/sphere.obj',
]


def make_ref(src='/root/reference', quiet=False):
    """Returns DST, or None when the reference tree is not present (then whatever oracle/_ref already holds is used)."""
    if not os.path.isdir(src):
        return None
    manifest = {}
    for rel in FILES:
        a, b = os.path.join(src, rel), os.path.join(DST, rel)
        if not os.path.exists(a):
            if not quiet:
                print('make_ref: missing in the reference tree: %s' % rel)
            continue
        os.makedirs(os.path.dirname(b), exist_ok=True)
        shutil.copyfile(a, b)
        with open(a, 'rb') as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': src, 'sha256': manifest}, f, indent=1, sort_keys=True)
    return DST


if __name__ == '__main__':
    src = sys.argv[sys.argv.index('--src') + 1] if '--src' in sys.argv else '/root/reference'
    print(make_ref(src))
