"""Generate golden input/output vectors by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py            # needs /root/reference (read-only), CPU only

The reference (``diffrend.torch.renderer.render``, renderer.py:136) is imported from
``/root/reference`` and run on torch-CPU; for every case the scene tensors, the kwargs, the outputs
(image/depth/normal/pos/nearest/ray_dir) and the autograd gradients of a seeded weighted loss are
written to ``tests/golden/<case>.npz``.  The fixtures travel to the GPU box; the reference does not.
Scene construction follows the reference's own call sites (cited per case).
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('SURF_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, REF)

from diffrend.torch.renderer import render as ref_render            # noqa: E402
from diffrend.torch.render import load_scene, make_torch_var         # noqa: E402
from diffrend.torch.params import SCENE_BASIC, SCENE_2                # noqa: E402
from diffrend.torch.utils import tch_var_f, tch_var_l                 # noqa: E402
from diffrend.model import load_model, obj_to_triangle_spec           # noqa: E402
from diffrend.utils.sample_generator import uniform_sample_mesh, uniform_sample_sphere  # noqa: E402

import scene_io                                                       # noqa: E402
from surf_renderer_b200 import scenes as synth                        # noqa: E402


def run_case(name, scene, params, loss_seed, with_grad=True, hit_only_geom=True, skip_leaves=()):
    sc = scene_io.clone_scene(scene, requires_grad=with_grad)
    leaves = scene_io.grad_leaves(sc) if with_grad else {}
    for k in skip_leaves:
        leaves.pop(k, None)
    res = ref_render(sc, **params)
    outs = {k: res[k] for k in ('image', 'depth', 'normal', 'pos', 'nearest', 'ray_dir')}
    grads = {}
    if with_grad:
        H, W = res['depth'].shape
        w = scene_io.loss_weights((H, W), loss_seed)
        loss = scene_io.weighted_loss(res, w, sc['camera']['far'], hit_only_geom=hit_only_geom)
        names = [k for k, v in leaves.items() if v.requires_grad]
        gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
        for k, g in zip(names, gs):
            if g is not None:
                grads[k] = g
    extra = {'loss_seed': loss_seed, 'hit_only_geom': hit_only_geom,
             'hit_pixels': int((res['depth'] <= sc['camera']['far']).sum())}
    path = os.path.join(HERE, name + '.npz')
    scene_io.save_case(path, scene, params, outs, grads, extra)
    print('%-28s M=%-6d N=%-6d hit=%-6d grads=%d  %.0f KB' % (
        name, sum(int(v['material_idx'].shape[0]) for v in scene['objects'].values()),
        res['depth'].numel(), extra['hit_pixels'], len(grads), os.path.getsize(path) / 1024))


def main():
    torch.manual_seed(0)
    # A: scenes/basic.json at 64x64 (torch/render.py:37-107 loader, BASELINE configs[0])
    sc = load_scene(os.path.join(REF, 'scenes/basic.json'))
    sc['camera']['viewport'] = [0, 0, 64, 64]
    sc = make_torch_var(sc)
    run_case('a_basic_json_64', sc, {}, 11)

    # A': plane + sphere + disk, one light; sphere grads are NaN in the reference -> skipped
    run_case('a_basic_mixed_64', synth.basic_mixed(64, 64), {}, 12,
             skip_leaves=('objects/sphere/pos', 'objects/sphere/radius'))

    # SCENE_BASIC as shipped (params.py:6-88) at 80x60
    sc = copy.deepcopy(SCENE_BASIC)
    sc['camera']['viewport'] = [0, 0, 80, 60]
    run_case('scene_basic_80x60', sc, {}, 13)
    run_case('scene_basic_80x60_ds_quartic', sc, {'double_sided': True, 'use_quartic': True}, 14)
    run_case('scene_basic_80x60_missgeom', sc, {}, 15, hit_only_geom=False)

    # B (reduced): bunny.splat exactly as test_scalability builds it (test_optimization.py:634-655) at 48x48
    splats = load_model(os.path.join(REF, 'data/bunny.splat'))
    v = splats['v']
    v = (v - np.mean(v, axis=0)) / (v.max() - v.min())
    sc = copy.deepcopy(SCENE_BASIC)
    sc['camera']['viewport'] = [0, 0, 48, 48]
    sc['camera']['fovy'] = np.deg2rad(5.)
    sc['camera']['focal_length'] = 2.
    sc['objects']['disk']['pos'] = tch_var_f(v)
    sc['objects']['disk']['normal'] = tch_var_f(splats['vn'])
    sc['objects']['disk']['radius'] = tch_var_f(splats['r'].ravel() * 2)
    sc['objects']['disk']['material_idx'] = tch_var_l(np.zeros(v.shape[0], dtype=int).tolist())
    sc['materials']['albedo'] = tch_var_f([[0.6, 0.6, 0.6]])
    sc['materials']['coeffs'] = tch_var_f([[0.5, 0.4, 8.0]])
    run_case('b_bunny_48', sc, {}, 16)

    # C (reduced): torus_1K triangles (batch_render.py:63-97) at 64x64, 3 lights, double sided
    obj = load_model(os.path.join(REF, 'data/torus_1K.obj'))
    vv = obj['v']
    vv = (vv - np.mean(vv, axis=0)) / max(np.max(vv, axis=0) - np.min(vv, axis=0))
    obj['v'] = vv
    mesh = obj_to_triangle_spec(obj)
    sc = copy.deepcopy(SCENE_BASIC)
    del sc['objects']['disk']
    sc['camera']['viewport'] = [0, 0, 64, 64]
    sc['camera']['fovy'] = np.deg2rad(18.)
    sc['camera']['focal_length'] = 0.1
    sc['camera']['eye'] = tch_var_f([3., 3., 3., 1.])
    sc['objects']['triangle'] = {'face': tch_var_f(mesh['face'].tolist()),
                                 'normal': tch_var_f(mesh['normal'].tolist()),
                                 'material_idx': tch_var_l(np.zeros(mesh['face'].shape[0], dtype=int).tolist())}
    sc['lights'] = {k: (val[:3] if k != 'ambient' else val) for k, val in sc['lights'].items()}
    sc['materials']['albedo'] = tch_var_f([[0.6, 0.6, 0.6]])
    sc['materials']['coeffs'] = tch_var_f([[0.5, 0.4, 8.0]])
    run_case('c_torus_64', sc, {'double_sided': True}, 17)

    # D (one element, reduced): 5000 chair splats (full_diff_renderer_demo.py:39-92,353-360) at 40x40
    np.random.seed(1000)
    obj = load_model(os.path.join(REF, 'data/chair_0001.off'))
    vv = obj['v']
    vv = (vv - np.mean(vv, axis=0)) / max(np.max(vv, axis=0) - np.min(vv, axis=0))
    obj['v'] = vv
    pv, pn = uniform_sample_mesh(obj, num_samples=5000)
    cam = uniform_sample_sphere(radius=5.0, num_samples=1)
    sc = copy.deepcopy(SCENE_BASIC)
    sc['camera']['viewport'] = [0, 0, 40, 40]
    sc['camera']['fovy'] = np.deg2rad(18.)
    sc['camera']['focal_length'] = 0.1
    sc['camera']['at'] = tch_var_f(np.mean(vv, axis=0))
    sc['camera']['eye'] = tch_var_f(cam[0])
    sc['objects']['disk']['pos'] = tch_var_f(pv)
    sc['objects']['disk']['normal'] = tch_var_f(pn)
    sc['objects']['disk']['radius'] = tch_var_f(np.ones(5000) * 0.025)
    sc['objects']['disk']['material_idx'] = tch_var_l(np.zeros(5000, dtype=int).tolist())
    sc['materials']['albedo'] = tch_var_f([[0.6, 0.6, 0.6]])
    sc['tonemap']['gamma'] = tch_var_f([1.0])
    run_case('d_chair_splats_40', sc, {'double_sided': True}, 18)

    # halfbox_sphere_cube.json (the scene projection_layer.py:461-846's consistency tests render) at 48x36
    sc = load_scene(os.path.join(REF, 'scenes/halfbox_sphere_cube.json'))
    sc['camera']['viewport'] = [0, 0, 48, 36]
    sc = make_torch_var(sc)
    run_case('halfbox_sphere_cube_48x36', sc, {}, 19)

    # SCENE_2 (params.py:177-258, orthographic, disks + spheres + triangles); coeffs added.
    # 60x45 = 2700 px <= tile_size so the reference's tiled ortho path works.
    sc = copy.deepcopy(SCENE_2)
    sc['camera']['viewport'] = [0, 0, 60, 45]
    sc['materials']['coeffs'] = tch_var_f([[1.0, 0.0, 0.0]] * 6)
    run_case('scene2_ortho_60x45', sc, {}, 20, with_grad=False)
    sc['camera']['proj_type'] = 'persp'
    sc['camera']['fovy'] = np.deg2rad(60.)
    sc['camera']['focal_length'] = 1.0
    run_case('scene2_persp_60x45', sc, {}, 21, with_grad=False)

    # orthographic camera WITH gradients: no spheres (their gradients are NaN in the reference), 56x40 <= one tile
    run_case('ortho_mixed_nosphere_56x40', synth.random_mixed_scene(8, width=56, height=40, n_sphere=0, n_disk=25, n_tri=20,
                                                                    proj='orthographic'), {'double_sided': True}, 27)

    # E (reduced): 3000 synthetic sphere-shell splats at 40x40 with the config-E camera
    run_case('e_synth_3000_40', synth.config_e(m=3000, width=40, height=40, radius=0.03), {}, 22)

    # random scenes with all four primitive types, Phong materials, attenuation, homogeneous coords
    run_case('mixed_r1', synth.random_mixed_scene(1), {}, 23, with_grad=False)
    run_case('mixed_r2_ds', synth.random_mixed_scene(2, homogeneous=True,
                                                     order=('triangle', 'plane', 'disk', 'sphere')),
             {'double_sided': True}, 24, with_grad=False)
    run_case('mixed_r3_nosphere', synth.random_mixed_scene(3, n_sphere=0, n_disk=30, n_tri=25),
             {'double_sided': True, 'use_quartic': True, 'tile_size': 500}, 25)
    run_case('mixed_r4_nosphere', synth.random_mixed_scene(4, n_sphere=0, n_plane=2, homogeneous=True,
                                                           order=('plane', 'triangle', 'disk')), {}, 26)


if __name__ == '__main__':
    main()
