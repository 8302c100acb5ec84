"""CPU tests of the boundary: the C-ABI shared library loads and exports every symbol include/surf_b200.h declares
(no compute calls), the ctypes mirror matches the header, and the host-side scene marshalling / error behaviour
mirrors the reference's render() (diffrend/torch/renderer.py:136-355)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import scene_io
from surf_renderer_b200 import _abi, scenes as synth
from surf_renderer_b200.marshal import Marshalled, make_options

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'surf_b200.h')


def _header_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return set(re.findall(r'\b(surf_[a-z_0-9]+)\s*\(', text))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from surf_renderer_b200 import build
    so = build.build()
    assert os.path.exists(so)
    handle = C.CDLL(so)
    declared = _header_functions()
    assert declared == set(_abi.SYMBOLS), (declared ^ set(_abi.SYMBOLS))
    for name in declared:
        assert hasattr(handle, name), name
    _abi.bind(handle)
    assert handle.surf_abi_version() == _abi.SURF_ABI_VERSION           # host-only call, no GPU needed
    assert handle.surf_workspace_bytes(1000, 4096, 3, 0) > 4096 * 20
    assert handle.surf_workspace_bytes(1000, 4096, 3, 1) >= handle.surf_workspace_bytes(1000, 4096, 3, 0) + 4096 * 12


def test_library_is_sm100a_with_tma_and_packed_fma():
    """The shipped cubin is sm_100a and the hot kernel contains TMA bulk copies, mbarrier waits and FFMA2."""
    import shutil
    import subprocess
    from surf_renderer_b200 import build
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', build.build()], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass
    for mnemonic in ('UBLKCP', 'SYNCS.PHASECHK', 'FFMA2', 'MUFU.RCP'):
        assert mnemonic in sass, mnemonic


def test_struct_layouts_match_c():
    """ctypes mirrors must have the C compiler's layout: check sizes against a tiny C program."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu", sizeof(SurfPrimSet), sizeof(SurfScene), sizeof(SurfCamera), sizeof(SurfOptions), sizeof(SurfOutputs), sizeof(SurfOutGrads), sizeof(SurfSceneGrads), sizeof(SurfBatchLayout), sizeof(SurfSplats), sizeof(SurfSplatGrads));return 0;}\n' % HEADER
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 'a.c')
        open(c, 'w').write(src)
        exe = os.path.join(d, 'a.out')
        subprocess.check_call(['gcc', '-std=c99', c, '-o', exe])        # the header is plain C
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mine = [C.sizeof(t) for t in (_abi.SurfPrimSet, _abi.SurfScene, _abi.SurfCamera, _abi.SurfOptions,
                                  _abi.SurfOutputs, _abi.SurfOutGrads, _abi.SurfSceneGrads, _abi.SurfBatchLayout,
                                  _abi.SurfSplats, _abi.SurfSplatGrads)]
    assert sizes == mine


def test_marshal_accepts_reference_style_inputs():
    scene = synth.random_mixed_scene(5, homogeneous=True)
    scene['camera']['viewport'] = torch.tensor([0, 0, 48, 40])          # make_torch_var turns lists into int64 tensors
    scene['objects']['disk']['material_idx'] = scene['objects']['disk']['material_idx'].float()   # may arrive as float
    scene['lights']['color_idx'] = scene['lights']['color_idx'].tolist()
    scene['colors'] = scene['colors'].numpy()
    scene['tonemap']['gamma'] = 0.8
    m = Marshalled(scene, 'cpu')
    assert (m.width, m.height) == (48, 40)
    assert [s[0] for s in m.sets] == list(scene['objects'].keys())
    assert m.sets[0][2] == 4 and m.sets[0][3] == 4                      # homogeneous strides kept, no copy
    sc = m.c_scene()
    assert sc.n_sets == 4 and sc.sets[0].kind == _abi.KIND['disk'] and sc.sets[0].pos_stride == 4
    assert m.ints['objects/disk/material_idx'].dtype == torch.int32
    assert sc.gamma is not None
    cam = m.c_camera()
    assert cam.proj == 0 and abs(cam.fovy - scene['camera']['fovy']) < 1e-12


def test_strided_batch_marshalling():
    """render_batch host logic: a list of same-shaped scenes stacks into one batched dict (shared objects stay
    shared), and MarshalledBatch turns it into scene-0 pointers + per-scene element strides without copying."""
    from surf_renderer_b200.marshal import MarshalledBatch
    from surf_renderer_b200.renderer import _stack_scenes
    shared_lights = torch.rand(3, 4)
    scenes = []
    for i in range(5):
        sc = synth.config_d_scene(i, m=40, width=16, height=12)
        sc['lights']['pos'] = shared_lights
        sc['lights']['attenuation'] = sc['lights']['attenuation'][:3]
        sc['lights']['color_idx'] = [1, 2, 3]                       # a python list, equal in every scene
        scenes.append(sc)
    stacked = _stack_scenes(scenes)
    assert stacked is not None
    assert stacked['objects']['disk']['pos'].shape == (5, 40, 3) and stacked['objects']['disk']['radius'].shape == (5, 40)
    assert stacked['lights']['pos'] is shared_lights
    assert stacked['camera']['eye'].shape[0] == 5 and stacked['tonemap']['gamma'].shape == (5, 1)
    assert torch.equal(stacked['objects']['disk']['pos'][3], scenes[3]['objects']['disk']['pos'])

    mb = MarshalledBatch(stacked, 'cpu')
    assert mb.batch == 5 and mb.m.total_prims == 40 and (mb.m.width, mb.m.height) == (16, 12)
    lay = mb.c_layout()
    assert lay.set_pos[0] == 120 and lay.set_normal[0] == 120 and lay.set_radius[0] == 40 and lay.set_material_idx[0] == 40
    assert lay.light_pos == 0 and lay.light_attenuation == 9 and lay.light_color_idx == 3
    assert lay.eye == stacked['camera']['eye'].shape[1] and lay.gamma == 1
    sc0 = mb.m.c_scene(mb.views0(mb.fulls))
    i_pos = mb.m.sets[0][4]['pos']
    assert sc0.sets[0].pos == mb.fulls[i_pos].data_ptr() and sc0.n_lights == 3
    assert mb.m.c_camera().eye == mb.m.cam_vecs['eye'].data_ptr()

    # sub-batches (scenes over ranks, dist.shard_scenes): batched leaves are indexed, shared ones passed through
    from surf_renderer_b200 import select_scenes
    from surf_renderer_b200.dist import shard_scenes
    idx = shard_scenes(5, 1, 2)
    part = select_scenes(stacked, idx)
    assert part['objects']['disk']['pos'].shape == (2, 40, 3) and part['lights']['pos'] is shared_lights
    assert torch.equal(part['camera']['eye'][1], stacked['camera']['eye'][3])
    assert part['camera']['viewport'] == stacked['camera']['viewport']
    assert MarshalledBatch(part, 'cpu').batch == 2

    # the same scene object repeated has nothing to stack along: render_batch then marshals scene by scene
    from surf_renderer_b200.marshal import batched_scene_size
    same = _stack_scenes([scenes[0], scenes[0], scenes[0]])
    assert same is not None and batched_scene_size(same) is None and batched_scene_size(stacked) == 5

    # different primitive counts, a different viewport or a different kind order cannot be stacked
    odd = synth.config_d_scene(9, m=41, width=16, height=12)
    assert _stack_scenes(scenes + [odd]) is None
    wide = synth.config_d_scene(9, m=40, width=20, height=12)
    assert _stack_scenes(scenes + [wide]) is None
    with pytest.raises(ValueError):
        MarshalledBatch(synth.config_d_scene(0, m=40, width=16, height=12), 'cpu')     # nothing batched
    bad = dict(stacked)
    bad['colors'] = torch.rand(4, 3, 3)
    with pytest.raises(ValueError):
        MarshalledBatch(bad, 'cpu')                                                      # 4 != 5


def test_marshal_errors_mirror_reference():
    scene = synth.scene_basic(16, 12)
    bad = scene_io.clone_scene(scene); bad['camera']['proj_type'] = 'fisheye'
    with pytest.raises(ValueError):
        Marshalled(bad, 'cpu')
    bad = scene_io.clone_scene(scene); del bad['lights']['ambient']
    with pytest.raises(KeyError):
        Marshalled(bad, 'cpu')                                          # renderer.py:274 requires it
    bad = scene_io.clone_scene(scene); del bad['materials']['coeffs']
    with pytest.raises(KeyError):
        Marshalled(bad, 'cpu')                                          # renderer.py:277
    bad = scene_io.clone_scene(scene); bad['objects'] = {'torus': bad['objects']['disk']}
    with pytest.raises(KeyError):
        Marshalled(bad, 'cpu')                                          # utils.py:488
    bad = scene_io.clone_scene(scene); bad['objects']['disk']['material_idx'] = torch.tensor([0, 1, 99])
    with pytest.raises(IndexError):
        Marshalled(bad, 'cpu')


def test_render_kwargs_behaviour_without_gpu():
    import surf_renderer_b200
    scene = synth.scene_basic(16, 12)
    with pytest.raises(RuntimeError):
        surf_renderer_b200.render(scene, vis_stat=True)                 # renderer.py:233-234
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match='no CPU path'):
            surf_renderer_b200.render(scene)                            # fails loudly, never falls back
    assert surf_renderer_b200.get_param_value('a', {'a': 3}, 5) == 3
    assert surf_renderer_b200.get_param_value('b', {'a': 3}, 5) == 5
    with pytest.raises(ValueError):
        surf_renderer_b200.get_param_value('b', {}, None, required=True)


def test_options_mapping():
    opt = make_options({'double_sided': True, 'use_quartic': 1, 'shadow': False, 'tile_size': 17}, (10, 20))
    assert (opt.double_sided, opt.use_quartic, opt.shadow, opt.pixel_begin, opt.pixel_end) == (1, 1, 0, 10, 20)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no file of the package (or bench's non-baseline path) may import it."""
    pkg = os.path.join(ROOT, 'surf_renderer_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text and 'torch_oracle' not in text, f


def test_workspace_sizes_cover_the_candidate_queue():
    """perspective frames above 256 x 256 pixels carry the candidate queue of the constant-bank intersection path (80 B per
    pixel + flags + the `inside` list); surf_workspace_bytes() bounds every camera; small frames and orthographic ones do
    not pay for it"""
    from surf_renderer_b200._lib import lib
    L = lib()
    prims, lights = 100000, 3
    for n in (64 * 64, 256 * 256):
        assert L.surf_workspace_bytes_ex(prims, n, lights, 0, 0, 0) <= L.surf_workspace_bytes_ex(prims, n, lights, 0, 1, 0)
    n = 1024 * 1024
    persp, ortho = L.surf_workspace_bytes_ex(prims, n, lights, 0, 0, 0), L.surf_workspace_bytes_ex(prims, n, lights, 0, 1, 0)
    w0, w1 = L.surf_workspace_bytes_ex(prims, 64 * 64, lights, 0, 0, 0), L.surf_workspace_bytes_ex(prims, 256 * 256, lights, 0, 0, 0)
    without_queue = w1 + (w1 - w0) / (256 * 256 - 64 * 64) * (n - 256 * 256)       # linear in the pixel count below the threshold
    assert 80 * n <= persp - without_queue <= 82 * n            # the queue: 10 entries of 8 bytes per pixel (+ flags, inside list)
    assert persp - ortho >= 40 * n - (1 << 20)                 # 80 B per pixel of queue against 40 B per pixel of generic rays
    assert L.surf_workspace_bytes(prims, n, lights, 0) == max(persp, ortho)
    assert L.surf_workspace_bytes_ex(prims, n, lights, 0, 0, 1) - persp >= 12 * n      # + d(loss)/d(image) of a fused step
