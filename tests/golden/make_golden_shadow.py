"""Golden vectors of the reference's SHADOW-RAY branch (renderer.py:291-314), run in this container on torch-CPU.

    python tests/golden/make_golden_shadow.py        # needs /root/reference

The branch casts its visibility with ``.type(torch.cuda.FloatTensor)`` (renderer.py:311), so the stock reference can
only take it on a CUDA build of torch.  Here the attribute ``torch.cuda.FloatTensor`` is pointed at
``torch.FloatTensor`` before the reference is imported - the reference's files are untouched and every other line of
the branch (per-light shadow rays from ``frag_pos + 0.1 L``, ``ray_object_intersections`` with per-pixel origins,
the ``0 < t < |L|`` window, the own-primitive exemption) runs as shipped.  Fixtures: tests/golden/sh_*.npz, same
format as make_golden.py (scene, kwargs incl. shadow=True, outputs, autograd gradients of a seeded weighted loss).
"""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

torch.cuda.FloatTensor = torch.FloatTensor        # see the module docstring

import make_golden as mg                          # noqa: E402  (imports the reference)
from make_golden import SCENE_BASIC, load_model, obj_to_triangle_spec, run_case, synth, tch_var_f, tch_var_l   # noqa: E402


def main():
    torch.manual_seed(0)
    sc = copy.deepcopy(SCENE_BASIC)
    sc['camera']['viewport'] = [0, 0, 64, 48]
    run_case('sh_scene_basic_64x48', sc, {'shadow': True}, 41)

    # bunny.splat as the demos render it (shadow=True is their default: full_diff_renderer_demo.py:378), 32x32
    splats = load_model(os.path.join(mg.REF, 'data/bunny.splat'))
    v = splats['v']
    v = (v - np.mean(v, axis=0)) / (v.max() - v.min())
    sc = copy.deepcopy(SCENE_BASIC)
    sc['camera']['viewport'] = [0, 0, 32, 32]
    sc['camera']['fovy'] = np.deg2rad(5.)
    sc['camera']['focal_length'] = 2.
    sc['objects']['disk']['pos'] = tch_var_f(v)
    sc['objects']['disk']['normal'] = tch_var_f(splats['vn'])
    sc['objects']['disk']['radius'] = tch_var_f(splats['r'].ravel() * 2)
    sc['objects']['disk']['material_idx'] = tch_var_l(np.zeros(v.shape[0], dtype=int).tolist())
    sc['materials']['albedo'] = tch_var_f([[0.6, 0.6, 0.6]])
    sc['materials']['coeffs'] = tch_var_f([[0.5, 0.4, 8.0]])
    sc['lights'] = {k: (val[:3] if k != 'ambient' else val) for k, val in sc['lights'].items()}
    run_case('sh_bunny_32', sc, {'shadow': True}, 42)

    # torus triangles, double sided, 3 lights
    obj = load_model(os.path.join(mg.REF, 'data/torus_1K.obj'))
    vv = obj['v']
    obj['v'] = (vv - np.mean(vv, axis=0)) / max(np.max(vv, axis=0) - np.min(vv, axis=0))
    mesh = obj_to_triangle_spec(obj)
    sc = copy.deepcopy(SCENE_BASIC)
    del sc['objects']['disk']
    sc['camera']['viewport'] = [0, 0, 40, 40]
    sc['camera']['fovy'] = np.deg2rad(18.)
    sc['camera']['focal_length'] = 0.1
    sc['camera']['eye'] = tch_var_f([3., 3., 3., 1.])
    sc['objects']['triangle'] = {'face': tch_var_f(mesh['face'].tolist()), 'normal': tch_var_f(mesh['normal'].tolist()),
                                 'material_idx': tch_var_l(np.zeros(mesh['face'].shape[0], dtype=int).tolist())}
    sc['lights'] = {k: (val[:3] if k != 'ambient' else val) for k, val in sc['lights'].items()}
    sc['materials']['albedo'] = tch_var_f([[0.6, 0.6, 0.6]])
    sc['materials']['coeffs'] = tch_var_f([[0.5, 0.4, 8.0]])
    run_case('sh_torus_40_ds', sc, {'shadow': True, 'double_sided': True}, 43)

    # synthetic shell splats (config E shape) and random scenes with every primitive kind
    run_case('sh_synth_1500_36', synth.config_e(m=1500, width=36, height=36, radius=0.05), {'shadow': True}, 44)
    run_case('sh_mixed_r5', synth.random_mixed_scene(5, width=44, height=32, n_disk=30, n_tri=20, n_sphere=4),
             {'shadow': True}, 45, with_grad=False)
    run_case('sh_mixed_r6_nosphere', synth.random_mixed_scene(6, n_sphere=0, n_disk=30, n_tri=25, n_plane=1),
             {'shadow': True, 'double_sided': True, 'tile_size': 300}, 46)


if __name__ == '__main__':
    main()
