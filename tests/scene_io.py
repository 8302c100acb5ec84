"""Scene-dict <-> flat ``.npz`` helpers shared by the golden-fixture generator and the tests.

The scene schema is the reference's (diffrend/torch/renderer.py:136-355 reads it).
Dict insertion order of ``scene['objects']`` matters (utils.py:486) and is stored explicitly.
"""
from __future__ import annotations

import json

import numpy as np
import torch

FLOAT_LEAVES = (
    'materials/albedo', 'materials/coeffs', 'lights/pos', 'lights/attenuation',
    'lights/ambient', 'colors', 'tonemap/gamma',
)
PRIM_FLOAT_FIELDS = {
    'disk': ('pos', 'normal', 'radius'),
    'plane': ('pos', 'normal'),
    'sphere': ('pos', 'radius'),
    'triangle': ('face', 'normal'),
}


def _walk(d, prefix=''):
    for k, v in d.items():
        key = prefix + k
        if isinstance(v, dict):
            yield from _walk(v, key + '/')
        else:
            yield key, v


def flatten_scene(scene):
    """scene dict -> ({key: np.ndarray}, meta dict).  Non-array scalars/strings go to meta."""
    arrays, meta = {}, {'objects_order': list(scene['objects'].keys()), 'scalars': {}}
    for key, v in _walk(scene):
        if isinstance(v, torch.Tensor):
            arrays[key] = v.detach().cpu().numpy()
        elif isinstance(v, np.ndarray):
            arrays[key] = v
        elif isinstance(v, (list, tuple)):
            arrays[key] = np.asarray(v)
            meta.setdefault('lists', []).append(key)
        elif isinstance(v, (np.floating, np.integer)):
            meta['scalars'][key] = v.item()
        else:
            meta['scalars'][key] = v
    return arrays, meta


def _set(d, path, value):
    parts = path.split('/')
    for p in parts[:-1]:
        d = d.setdefault(p, {})
    d[parts[-1]] = value


def unflatten_scene(arrays, meta, device='cpu'):
    """Inverse of :func:`flatten_scene`; float arrays become f32 tensors, integer ones int64 tensors."""
    scene = {}
    # objects first, in the recorded order
    scene['objects'] = {k: {} for k in meta['objects_order']}
    lists = set(meta.get('lists', []))
    for key, arr in arrays.items():
        if key in lists and key == 'camera/viewport':
            _set(scene, key, [int(x) for x in arr.tolist()])
            continue
        if np.issubdtype(arr.dtype, np.floating):
            t = torch.tensor(arr, dtype=torch.float32, device=device)
        else:
            t = torch.tensor(arr, dtype=torch.int64, device=device)
        _set(scene, key, t)
    for key, v in meta['scalars'].items():
        _set(scene, key, v)
    return scene


def save_case(path, scene, params, outputs, grads=None, extra=None):
    arrays, meta = flatten_scene(scene)
    blob = {'scene/' + k: v for k, v in arrays.items()}
    for k, v in outputs.items():
        blob['out/' + k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    for k, v in (grads or {}).items():
        blob['grad/' + k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    meta['params'] = params
    meta['extra'] = extra or {}
    blob['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **blob)


def load_case(path, device='cpu'):
    z = np.load(path)
    meta = json.loads(bytes(z['meta']).decode())
    arrays = {k[len('scene/'):]: z[k] for k in z.files if k.startswith('scene/')}
    scene = unflatten_scene(arrays, meta, device=device)
    outs = {k[len('out/'):]: z[k] for k in z.files if k.startswith('out/')}
    grads = {k[len('grad/'):]: z[k] for k in z.files if k.startswith('grad/')}
    return scene, meta['params'], outs, grads, meta.get('extra', {})


def grad_leaves(scene):
    """Ordered {name: tensor} of the float leaves the render path is differentiable in."""
    leaves = {}
    for kind, prim in scene['objects'].items():
        for f in PRIM_FLOAT_FIELDS[kind]:
            leaves['objects/%s/%s' % (kind, f)] = prim[f]
    for name in FLOAT_LEAVES:
        a, b = name.split('/') if '/' in name else (name, None)
        if a not in scene:
            continue
        v = scene[a] if b is None else scene[a].get(b)
        if isinstance(v, torch.Tensor) and v.is_floating_point():
            leaves[name] = v
    return leaves


from surf_renderer_b200.scenes import clone_scene   # noqa: E402,F401  (kept here for the tests' convenience)


def loss_weights(shape_hw, seed):
    """Deterministic random loss weights for image/depth/pos/normal (CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    H, W = shape_hw
    return {
        'image': torch.rand(H, W, 3, generator=g) - 0.3,
        'depth': torch.rand(H, W, generator=g) - 0.5,
        'pos': torch.rand(H, W, 3, generator=g) - 0.5,
        'normal': torch.rand(H, W, 3, generator=g) - 0.5,
    }


def weighted_loss(res, weights, far, hit_only_geom=True):
    """Scalar loss touching every differentiable output.  pos/normal are weighted on hit pixels only by
    default (the reference back-propagates miss-pixel pos/normal into primitive 0 through an
    ill-conditioned far plane hit - SURVEY appendix B, miss-pixel note)."""
    dev = res['image'].device
    w = {k: v.to(dev) for k, v in weights.items()}
    hit = (res['depth'].detach() <= far).float()
    geom_mask = hit[..., None] if hit_only_geom else 1.0
    return ((res['image'] * w['image']).sum() + (res['depth'] * w['depth'] * hit).sum()
            + (res['pos'] * w['pos'] * geom_mask).sum() + (res['normal'] * w['normal'] * geom_mask).sum())
