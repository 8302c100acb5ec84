"""The reference's four executable pins on render() geometry, re-pointed at the B200 renderer.

diffrend/torch/projection_layer.py holds the only assertions in the reference that consume render() output
(SURVEY section 4): test_raster_coordinates :428, test_render_projection_consistency :461,
test_transformation_consistency :609, test_depth_to_world_consistency :846, all on
scenes/halfbox_sphere_cube.json at its own 160x120 viewport.  The reference lives only in the build container
(no GPU) and the CUDA path only runs on the GPU box, so the check runs in two stages:

  GPU box :  python tools/projection_consistency.py export     -> gpurun_out/b200_halfbox_160x120.npz
             (renders the scene stored in tests/golden/halfbox_sphere_cube_48x36.npz at 160x120 through the C ABI)
  here    :  python tools/projection_consistency.py check FILE  (imports /root/reference, swaps the B200 outputs in
             for projection_layer.render_scene and runs the four reference tests unmodified)

The exported file is committed as tests/golden/b200_halfbox_160x120.npz; tests/test_projection_pins.py re-runs
`check` on it whenever /root/reference is present, and the GPU suite asserts that a fresh render still equals it.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

KEYS = ('image', 'depth', 'pos', 'normal', 'nearest')
VIEWPORT = [0, 0, 160, 120]            # scenes/halfbox_sphere_cube.json "viewport"
REF_SCENE = '/root/reference/scenes/halfbox_sphere_cube.json'


def export(path):
    import scene_io
    import surf_renderer_b200
    scene, _, _, _, _ = scene_io.load_case(os.path.join(ROOT, 'tests', 'golden', 'halfbox_sphere_cube_48x36.npz'))
    scene['camera']['viewport'] = list(VIEWPORT)
    res = surf_renderer_b200.render(scene_io.clone_scene(scene, device='cuda'))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **{k: res[k].detach().cpu().numpy() for k in KEYS})
    print('wrote', path, {k: tuple(res[k].shape) for k in KEYS})


def check(path, batch_size=6, verbose=True):       # 6 = the batch size of the reference's own main() (:895)
    """Run the four reference tests with the stored B200 outputs standing in for render_scene()."""
    if '/root/reference' not in sys.path:
        sys.path.insert(0, '/root/reference')
    import diffrend.torch.projection_layer as pl
    z = np.load(path)
    outs = {k: torch.tensor(z[k]) for k in KEYS}
    calls = []

    def b200_render_scene(scene_file):
        calls.append(scene_file)
        return {k: v.clone() for k, v in outs.items()}

    saved = pl.render_scene
    pl.render_scene = b200_render_scene
    ran = []
    try:
        for name in ('test_raster_coordinates', 'test_render_projection_consistency',
                     'test_transformation_consistency', 'test_depth_to_world_consistency'):
            getattr(pl, name)(REF_SCENE, batch_size)
            ran.append(name)
            if verbose:
                print('PASS', name)
    finally:
        pl.render_scene = saved
    assert len(calls) == 4
    return ran


if __name__ == '__main__':
    stage = sys.argv[1] if len(sys.argv) > 1 else 'check'
    default = os.path.join(ROOT, 'gpurun_out', 'b200_halfbox_160x120.npz') if stage == 'export' \
        else os.path.join(ROOT, 'tests', 'golden', 'b200_halfbox_160x120.npz')
    target = sys.argv[2] if len(sys.argv) > 2 else default
    if stage == 'export':
        export(target)
    else:
        check(target)
