"""Generates tests/golden/np_*.npz: outputs of the REFERENCE numpy twin renderer (diffrend/numpy/renderer.py::render)
on seeded scenes, used to pin oracle/numpy_oracle.py.  Runs only in the build container (imports /root/reference).

    python tests/golden/make_golden_numpy.py
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT_KEYS = ('image', 'depth', 'nearest', 'ray_dir')


# ---- flat storage of a numpy-twin scene -----------------------------------------------------------------------
def save_np_case(path, scene, outs):
    blob, meta = {}, {'objects_order': list(scene['objects'].keys()), 'scalars': {}, 'lists': {}}

    def walk(d, prefix):
        for k, v in d.items():
            key = prefix + k
            if isinstance(v, dict):
                walk(v, key + '/')
            elif isinstance(v, np.ndarray):
                blob['scene/' + key] = v
            elif isinstance(v, list):
                meta['lists'][key] = v
            else:
                meta['scalars'][key] = v
    walk(scene, '')
    for k in OUT_KEYS:
        blob['out/' + k] = np.asarray(outs[k])
    blob['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **blob)


def load_np_case(path):
    z = np.load(path)
    meta = json.loads(bytes(z['meta']).decode())
    scene = {'objects': {k: {} for k in meta['objects_order']}}

    def put(key, v):
        d = scene
        parts = key.split('/')
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    for k in z.files:
        if k.startswith('scene/'):
            put(k[6:], z[k])
    for k, v in meta['lists'].items():
        put(k, v)
    for k, v in meta['scalars'].items():
        put(k, v)
    return scene, {k[4:]: z[k] for k in z.files if k.startswith('out/')}


# ---- seeded scenes in the numpy twin's schema ------------------------------------------------------------------
def demo_like_scene(seed, W, H, list_camera):
    """disk + sphere + triangle (+ plane) with integer radii and python-list camera vectors, like the schema the
    reference's own __main__ block builds (renderer.py:304-356)."""
    r = np.random.RandomState(seed)

    def pts(n, lo, hi):
        return np.concatenate((r.uniform(lo, hi, (n, 3)), np.ones((n, 1))), axis=1)

    def dirs(n):
        return np.concatenate((r.normal(size=(n, 3)), np.zeros((n, 1))), axis=1)
    face = np.concatenate((r.uniform(-12, 12, (5, 3, 3)), np.ones((5, 3, 1))), axis=2)
    face[..., 2] -= 8
    fn = np.cross(face[:, 1, :3] - face[:, 0, :3], face[:, 2, :3] - face[:, 0, :3])
    fn = np.concatenate((fn, np.zeros((5, 1))), axis=1)
    eye, at, up = [0.0, 1.0, 10.0, 1.0], [0.0, 0.0, 0.0, 1.0], [0.0, 1.0, 0.0, 0.0]
    if not list_camera:
        eye, at, up = np.array([3.0, 2.0, 9.0, 1.0]), np.array([0.5, 0.0, 0.0, 1.0]), np.array([0.1, 1.0, 0.0, 0.0])
    scene = {
        'camera': {'viewport': [0, 0, W, H], 'fovy': float(np.deg2rad(70.)), 'focal_length': 1., 'eye': eye, 'up': up,
                   'at': at, 'near': 1.0, 'far': 1000.0},
        'lights': {'pos': np.array([[20., 20., 20., 1.0], [-15, 3., 15., 1.0], [2., -20., 5., 1.]]),
                   'color_idx': np.array([2, 1, 3]),
                   'attenuation': np.array([[0., 1., 0.], [0., 0., 1.], [1., 0., 0.]])},
        'colors': np.array([[0.0, 0.0, 0.0], [0.8, 0.1, 0.1], [0.2, 0.2, 0.2], [0.1, 0.6, 0.3]]),
        'materials': {'albedo': r.uniform(0.05, 0.95, (6, 3))},
        'objects': {
            'disk': {'normal': dirs(6), 'pos': pts(6, -6, 6), 'radius': np.array([4, 3, 2, 5, 1, 2]),
                     'material_idx': r.randint(0, 6, 6)},
            'sphere': {'pos': pts(3, -8, 8), 'radius': np.array([3.0, 2.0, 1.5]), 'material_idx': r.randint(0, 6, 3)},
            'triangle': {'face': face, 'normal': fn, 'material_idx': r.randint(0, 6, 5)},
        },
        'tonemap': {'type': 'gamma', 'gamma': 0.8},
    }
    if seed % 2:
        scene['objects']['plane'] = {'pos': np.array([[0., -9., 0., 1.]]), 'normal': np.array([[0., 1., 0.1, 0.]]),
                                     'material_idx': np.array([3])}
    return scene


def cases():
    sys.path.insert(0, ROOT)
    from oracle.numpy_oracle import homogeneous_scene
    from surf_renderer_b200 import scenes as synth
    yield 'np_demo_list_camera_64x48', demo_like_scene(2, 64, 48, True)
    yield 'np_demo_plane_array_camera_50x60', demo_like_scene(3, 50, 60, False)
    yield 'np_basic_mixed_64', homogeneous_scene(synth.basic_mixed(64, 64))
    yield 'np_synth_splats_2000_40', homogeneous_scene(synth.config_e(m=2000, width=40, height=40, radius=0.03))
    yield 'np_random_mixed_48x36', homogeneous_scene(
        synth.random_mixed_scene(11, width=48, height=36, n_disk=40, n_tri=30, n_sphere=5))


def main():
    sys.path.insert(0, '/root/reference')
    import copy
    from diffrend.numpy.renderer import render as ref_render
    for name, scene in cases():
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
            res = ref_render(copy.deepcopy(scene))
        save_np_case(os.path.join(HERE, name + '.npz'), scene, res)
        hit = np.isfinite(res['depth'])
        print('%-36s hit %.2f  winners %d  image max %.3f  nan %d' % (
            name, hit.mean(), len(np.unique(res['nearest'][hit])), np.nanmax(res['image']), int(np.isnan(res['image']).sum())))


if __name__ == '__main__':
    main()
