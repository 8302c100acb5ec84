"""In-code scene dictionaries (the role of the reference's ``diffrend/torch/params.py``) plus the
synthetic workloads BASELINE.json / SURVEY 8(d) name.  Pure data builders: CPU torch tensors in the
reference's scene-dict schema; nothing here touches the GPU or the reference tree.
"""
from __future__ import annotations

import copy
import math

import numpy as np
import torch


def _f(x):
    return torch.tensor(np.asarray(x, dtype=np.float64), dtype=torch.float32)


def _l(x):
    return torch.tensor(np.asarray(x), dtype=torch.int64)


def basic_lights_and_materials(n_lights=7, coeffs=(1.0, 0.0, 0.0)):
    """Values of SCENE_BASIC (diffrend/torch/params.py:6-88): 7 lights, 8 colours, 6 materials."""
    pos = [[10., 0., 0., 1.], [-10., 0., 0., 1.], [0., 10., 0., 1.], [0., -10., 0., 1.],
           [0., 0., 10., 1.], [0., 0., -10., 1.], [20., 20., 20., 1.]]
    cidx = [1, 3, 4, 5, 6, 7, 1]
    return {
        'lights': {
            'pos': _f(pos[:n_lights]),
            'color_idx': _l(cidx[:n_lights]),
            'attenuation': _f([[1., 0., 0.]] * n_lights),
            'ambient': _f([0.01, 0.01, 0.01]),
        },
        'colors': _f([[0.0, 0.0, 0.0], [0.8, 0.1, 0.1], [0.2, 0.2, 0.2], [0.2, 0.8, 0.2],
                      [0.2, 0.2, 0.8], [0.8, 0.2, 0.8], [0.8, 0.8, 0.2], [0.2, 0.8, 0.8]]),
        'materials': {
            'albedo': _f([[0.0, 0.0, 0.0], [0.1, 0.1, 0.1], [0.2, 0.2, 0.2], [0.5, 0.5, 0.5],
                          [0.9, 0.1, 0.1], [0.1, 0.6, 0.8]]),
            'coeffs': _f([list(coeffs)] * 6),
        },
        'tonemap': {'type': 'gamma', 'gamma': _f([0.8])},
    }


def basic_camera(width, height, fovy_deg=90., focal=1., eye=(0., 1., 10., 1.), at=(0., 0., 0., 1.),
                 near=0.1, far=1000., proj='perspective'):
    return {
        'proj_type': proj,
        'viewport': [0, 0, int(width), int(height)],
        'fovy': float(np.deg2rad(fovy_deg)),
        'focal_length': float(focal),
        'eye': _f(eye), 'up': _f([0., 1., 0., 0.]), 'at': _f(at),
        'near': float(near), 'far': float(far),
    }


def scene_basic(width=320, height=240):
    """SCENE_BASIC (params.py:6-88): three large disks, 7 lights."""
    s = basic_lights_and_materials()
    s['camera'] = basic_camera(width, height)
    s['objects'] = {'disk': {
        'normal': _f([[0., 0., 1., 0.], [0., 1., 0., 0.], [-1., -1., 1., 0.]]),
        'pos': _f([[0., -1., 3., 1.], [0., -1., 0., 1.], [10., 5., -5., 1.]]),
        'radius': _f([4, 7, 4]),
        'material_idx': _l([4, 3, 5]),
    }}
    return s


def basic_mixed(width=64, height=64):
    """Config A' (SURVEY 8d): plane + sphere + disk, one point light, SCENE_BASIC camera/materials."""
    s = basic_lights_and_materials()
    s['lights'] = {'pos': _f([[20., 20., 20., 1.]]), 'color_idx': _l([1]),
                   'attenuation': _f([[1., 0., 0.]]), 'ambient': _f([0.01, 0.01, 0.01])}
    s['camera'] = basic_camera(width, height)
    s['objects'] = {
        'plane': {'pos': _f([[0., -3., 0., 1.]]), 'normal': _f([[0., 1., 0., 0.]]), 'material_idx': _l([3])},
        'sphere': {'pos': _f([[-2., 0., 0., 1.]]), 'radius': _f([1.5]), 'material_idx': _l([4])},
        'disk': {'pos': _f([[2., 0., 1., 1.]]), 'normal': _f([[0., .3, 1., 0.]]), 'radius': _f([1.5]),
                 'material_idx': _l([5])},
    }
    return s


def splat_scene(pos, normal, radius, width, height, fovy_deg, focal, eye, n_lights=7,
                coeffs=(0.5, 0.4, 8.0), albedo=(0.6, 0.6, 0.6), gamma=0.8, at=(0., 0., 0., 1.)):
    """Disk-splat scene in the shape ``test_scalability`` / the demos build (test_optimization.py:634-655)."""
    s = basic_lights_and_materials(n_lights)
    s['camera'] = basic_camera(width, height, fovy_deg, focal, eye, at=at)
    m = pos.shape[0]
    s['objects'] = {'disk': {'pos': _f(pos), 'normal': _f(normal),
                             'radius': _f(np.broadcast_to(np.asarray(radius, dtype=np.float64), (m,))),
                             'material_idx': _l(np.zeros(m, dtype=np.int64))}}
    s['materials'] = {'albedo': _f([list(albedo)]), 'coeffs': _f([list(coeffs)])}
    s['tonemap'] = {'type': 'gamma', 'gamma': _f([gamma])}
    return s


def synthetic_sphere_splats(m, seed=0, shell_radius=0.5, normal_jitter=0.05):
    """Config E geometry: m splats on a sphere shell, outward-ish normals."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(m, 3, generator=g)
    d = d / d.norm(dim=1, keepdim=True)
    pos = shell_radius * d
    nrm = d + normal_jitter * torch.randn(m, 3, generator=g)
    nrm = nrm / nrm.norm(dim=1, keepdim=True)
    return pos.numpy().astype(np.float64), nrm.numpy().astype(np.float64)


def config_e(m=100_000, width=1024, height=1024, seed=0, radius=0.005):
    """Inverse-rendering step workload: 100K synthetic splats at 1024x1024 (BASELINE.json configs[4])."""
    pos, nrm = synthetic_sphere_splats(m, seed)
    s = splat_scene(pos, nrm, radius, width, height, fovy_deg=14., focal=1., eye=(0., 0., 5., 1.),
                    n_lights=3, coeffs=(0.5, 0.4, 8.0), albedo=(0.6, 0.6, 0.6), gamma=0.8)
    return s


def config_e_target_scene(scene, seed=1, jitter=0.002):
    """Target of the inverse-rendering step: the same scene with splat centres jittered."""
    t = copy.deepcopy(scene)
    g = torch.Generator().manual_seed(seed)
    p = t['objects']['disk']['pos']
    t['objects']['disk']['pos'] = p + jitter * torch.randn(p.shape, generator=g)
    return t


def config_d_scene(index, m=5000, width=128, height=128, radius=0.025, cam_dist=5.0):
    """GAN splat batch element, mesh-free fallback of SURVEY 8(d): random splats on a 0.5 shell, camera on a
    radius-5 sphere looking at the origin, fovy 18 deg, focal 0.1, gamma 1, double-sided at render time."""
    seed = 1000 + index
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(m, 3, generator=g)
    d = d / d.norm(dim=1, keepdim=True)
    pos = 0.5 * d
    nrm = d + 0.1 * torch.randn(m, 3, generator=g)
    e = torch.randn(3, generator=g)
    e = cam_dist * e / e.norm()
    eye = (float(e[0]), float(e[1]), float(e[2]), 1.0)
    return splat_scene(pos.numpy(), nrm.numpy(), radius, width, height, fovy_deg=18., focal=0.1, eye=eye,
                       n_lights=7, coeffs=(1.0, 0.0, 0.0), gamma=1.0)


def config_d_batch(n_scenes=64, m=5000, width=128, height=128, radius=0.025, pin=False):
    """BASELINE configs[3] as ONE batched scene dict (render_batch's stacked form): splat positions / normals [B,M,3]
    and camera eyes [B,4] carry the batch dimension, everything else (radius, lights, materials) is shared.
    pin=True: the batched tensors live in pinned host memory (the per-step upload of an end-to-end measurement)."""
    parts = [config_d_scene(i, m=m, width=width, height=height, radius=radius) for i in range(n_scenes)]
    s = parts[0]
    disk = s['objects']['disk']
    disk['pos'] = torch.stack([p['objects']['disk']['pos'] for p in parts], 0).contiguous()
    disk['normal'] = torch.stack([p['objects']['disk']['normal'] for p in parts], 0).contiguous()
    s['camera']['eye'] = torch.stack([torch.as_tensor(p['camera']['eye'], dtype=torch.float32) for p in parts], 0).contiguous()
    if pin:
        disk['pos'], disk['normal'] = disk['pos'].pin_memory(), disk['normal'].pin_memory()
        s['camera']['eye'] = s['camera']['eye'].pin_memory()
    return s


def random_mixed_scene(seed, width=48, height=40, n_disk=12, n_plane=1, n_sphere=3, n_tri=10,
                       n_lights=3, n_mat=5, order=('disk', 'sphere', 'triangle', 'plane'),
                       homogeneous=False, proj='perspective'):
    """Random scene exercising all four primitive types, mixed Phong coefficients and attenuations."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, lo=-1.0, hi=1.0):
        return lo + (hi - lo) * torch.rand(*shape, generator=g)

    def homog(x, w):
        if not homogeneous:
            return x
        return torch.cat((x, torch.full(x.shape[:-1] + (1,), float(w))), dim=-1)

    objs = {}
    for kind in order:
        if kind == 'disk' and n_disk:
            objs['disk'] = {'pos': homog(rnd(n_disk, 3, lo=-2.5, hi=2.5), 1),
                            'normal': homog(rnd(n_disk, 3) + torch.tensor([0., 0., 1.2]), 0),
                            'radius': rnd(n_disk, lo=0.3, hi=1.2),
                            'material_idx': torch.randint(0, n_mat, (n_disk,), generator=g)}
        elif kind == 'plane' and n_plane:
            objs['plane'] = {'pos': homog(torch.tensor([[0., -3.5, 0.]]).repeat(n_plane, 1) + 0.1 * rnd(n_plane, 3), 1),
                             'normal': homog(torch.tensor([[0.05, 1., 0.1]]).repeat(n_plane, 1), 0),
                             'material_idx': torch.randint(0, n_mat, (n_plane,), generator=g)}
        elif kind == 'sphere' and n_sphere:
            objs['sphere'] = {'pos': homog(rnd(n_sphere, 3, lo=-3., hi=3.), 1),
                              'radius': rnd(n_sphere, lo=0.4, hi=1.1),
                              'material_idx': torch.randint(0, n_mat, (n_sphere,), generator=g)}
        elif kind == 'triangle' and n_tri:
            c = rnd(n_tri, 1, 3, lo=-3., hi=3.)
            v = c + rnd(n_tri, 3, 3, lo=-1.2, hi=1.2)
            nrm = torch.linalg.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
            nrm = nrm / nrm.norm(dim=1, keepdim=True)
            objs['triangle'] = {'face': homog(v, 1), 'normal': homog(nrm, 0),
                                'material_idx': torch.randint(0, n_mat, (n_tri,), generator=g)}
    s = {
        'camera': basic_camera(width, height, fovy_deg=50., focal=1., eye=(0.3, 0.8, 9.0, 1.0),
                               near=0.1, far=1000., proj=proj),
        'lights': {'pos': homog(rnd(n_lights, 3, lo=-12., hi=12.) + torch.tensor([0., 4., 8.]), 1),
                   'color_idx': torch.randint(0, 4, (n_lights,), generator=g),
                   'attenuation': torch.stack((rnd(n_lights, lo=0.6, hi=1.2), rnd(n_lights, lo=0.0, hi=0.05),
                                               rnd(n_lights, lo=0.0, hi=0.01)), dim=1),
                   'ambient': rnd(3, lo=0.0, hi=0.05)},
        'colors': rnd(4, 3, lo=0.1, hi=1.0),
        'materials': {'albedo': rnd(n_mat, 3, lo=0.1, hi=0.9),
                      'coeffs': torch.stack((rnd(n_mat, lo=0.3, hi=1.0), rnd(n_mat, lo=0.0, hi=0.6),
                                             torch.randint(1, 12, (n_mat,), generator=g).float()), dim=1)},
        'objects': objs,
        'tonemap': {'type': 'gamma', 'gamma': _f([0.8])},
    }
    if proj in ('ortho', 'orthographic'):
        s['camera']['fovy'] = float(np.deg2rad(120.))
        s['camera']['focal_length'] = 3.0
    return s


def clone_scene(scene, device=None, requires_grad=False):
    """Deep copy; tensors are detached clones (optionally moved / made leaves requiring grad)."""
    def rec(v):
        if isinstance(v, dict):
            return {k: rec(x) for k, x in v.items()}
        if isinstance(v, torch.Tensor):
            t = v.detach().clone()
            if device is not None:
                t = t.to(device)
            if requires_grad and t.is_floating_point():
                t.requires_grad_(True)
            return t
        if isinstance(v, list):
            return list(v)
        return v
    out = rec(scene)
    if requires_grad:
        # the reference cannot differentiate w.r.t. the camera (in-place op, utils.py:476)
        for k in ('eye', 'at', 'up'):
            if isinstance(out['camera'].get(k), torch.Tensor):
                out['camera'][k] = out['camera'][k].detach()
    return out
