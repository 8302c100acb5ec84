"""Scene / asset ingest (surf_renderer_b200/ingest.py) - SURVEY 8f-3.  Format semantics follow
diffrend/model.py:90-211 and diffrend/torch/render.py:9-107; when the reference tree is present the loaders are
compared against the reference's own on the reference's own assets."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from surf_renderer_b200 import ingest

REF = '/root/reference'
HAVE_REF = os.path.isdir(os.path.join(REF, 'diffrend'))


def _write(path, text):
    with open(path, 'w') as fh:
        fh.write(text)
    return str(path)


def test_obj_off_splat_parsing(tmp_path):
    obj = ingest.load_obj(_write(tmp_path / 'a.obj', '# c\nv 0 0 0\nv  1 0 0\nv 0 1 0\nvn 0 0 1\nv 0 0 1\nf 1/1/1 2/2/1 3/3/1\nf 1 3 4\n'))
    assert obj['v'].shape == (4, 3) and obj['f'].tolist() == [[0, 1, 2], [0, 2, 3]]
    off = ingest.load_off(_write(tmp_path / 'a.off', 'OFF\n4 2 0\n0 0 0\n1 0 0\n0 1 0\n0 0 1\n3 0 1 2\n3 0 2 3\n'))
    assert off['v'].shape == (4, 3) and off['f'].tolist() == [[0, 1, 2], [0, 2, 3]]
    off2 = ingest.load_off(_write(tmp_path / 'b.off', 'OFF 4 2 0\n0 0 0\n1 0 0\n0 1 0\n0 0 1\n3 0 1 2\n3 0 2 3\n'))
    assert np.array_equal(off2['v'], off['v']) and np.array_equal(off2['f'], off['f'])
    sp = ingest.load_splat(_write(tmp_path / 'a.splat', 'v 1 2 3\nvn 0 0 1\nr 0.5\nv 4 5 6\nvn 0 1 0\nr 0.25\n'))
    assert sp['v'].shape == (2, 3) and sp['r'].shape == (2, 1) and sp['type'] == 'splat'
    assert ingest.load_model(str(tmp_path / 'a.splat'))['vn'].tolist() == [[0, 0, 1], [0, 1, 0]]
    with pytest.raises(KeyError):
        ingest.load_model(str(tmp_path / 'a.ply'))


def test_triangle_spec_and_transform(tmp_path):
    obj = {'v': np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 0]]), 'f': np.array([[0, 1, 2], [0, 3, 3]])}
    spec = ingest.obj_to_triangle_spec(obj)
    assert spec['face'].shape == (2, 3, 4) and (spec['face'][..., 3] == 1).all()
    assert spec['normal'][0].tolist() == [0, 0, 1, 0]
    assert spec['normal'][1].tolist() == [0, 0, 0, 0]                      # degenerate face: zero normal kept
    M = ingest.axis_angle_matrix([0, 0, 2], np.pi / 2)
    assert np.allclose(M[:3, :3] @ np.array([1., 0, 0]), [0, 1, 0])
    t = ingest.transform_model({'v': np.array([[1., 0, 0]])}, [2, 2, 2], {'axis': [0, 0, 1], 'angle_deg': 90.}, [0, 0, 5])
    assert np.allclose(t['v'], [[0, 2, 5]])


def test_load_scene_and_make_torch_var(tmp_path):
    _write(tmp_path / 'tri.obj', 'v -1 -1 0\nv 1 -1 0\nv 0 1 0\nf 1 2 3\n')
    scene_json = {
        'camera': {'proj_type': 'perspective', 'viewport': [0, 0, 16, 12], 'fovy': 1.2, 'focal_length': 1.0,
                   'eye': [0.0, 0.0, 3.0, 1.0], 'up': [0.0, 1.0, 0.0, 0.0], 'at': [0.0, 0.0, 0.0, 1.0], 'near': 0.1, 'far': 1000.0},
        'lights': {'pos': [[2.0, 2.0, 4.0, 1.0]], 'color_idx': [1], 'attenuation': [[1.0, 0.0, 0.0]], 'ambient': [0.01, 0.01, 0.01]},
        'colors': [[0.0, 0.0, 0.0], [0.8, 0.8, 0.8]],
        'materials': {'albedo': [[0.5, 0.5, 0.5]], 'coeffs': [[1.0, 0.0, 0.0]]},
        'objects': {'obj': [{'path': './tri.obj', 'material_idx': 0},
                            {'path': './tri.obj', 'material_idx': 0, 'scale': [0.5, 0.5, 0.5], 'translate': [0.0, 0.0, 1.0]}]},
        'tonemap': {'type': 'gamma', 'gamma': [0.8]},
    }
    path = _write(tmp_path / 'scene.json', json.dumps(scene_json))
    scene = ingest.load_scene(path)
    assert 'obj' not in scene['objects'] and scene['objects']['triangle']['face'].shape == (2, 3, 4)
    assert np.allclose(scene['objects']['triangle']['face'][1, 2], [0, 0.5, 1.0, 1.0])
    sc = ingest.make_torch_var(scene, device='cpu')
    assert sc['camera']['viewport'].dtype == torch.int64                   # flat int list -> LongTensor
    assert sc['lights']['color_idx'].dtype == torch.int64
    assert sc['objects']['triangle']['material_idx'].dtype == torch.float32  # np.ones(...) * idx arrives as float
    assert sc['objects']['triangle']['face'].dtype == torch.float32 and sc['camera']['fovy'] == 1.2
    # the oracle (reference semantics) accepts the ingested scene as is
    from oracle import torch_oracle
    res = torch_oracle.render(sc)
    assert res['image'].shape == (12, 16, 3) and float(res['image'].max()) > 0


@pytest.mark.skipif(not HAVE_REF, reason='reference tree not present')
def test_loaders_match_reference_on_reference_assets():
    sys.path.insert(0, REF)
    from diffrend import model as rmodel
    from diffrend.torch import render as rrender
    for rel in ('data/torus_1K.obj', 'data/chair_0001.off', 'data/bunny.splat', 'scenes/objs/halfbox.obj'):
        a, b = ingest.load_model(os.path.join(REF, rel)), rmodel.load_model(os.path.join(REF, rel))
        for k in b:
            if k == 'type':
                assert a[k] == b[k]
            else:
                assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (rel, k)
    obj = rmodel.load_obj(os.path.join(REF, 'data/torus_1K.obj'), verbose=False)
    sa, sb = ingest.obj_to_triangle_spec(obj), rmodel.obj_to_triangle_spec(obj)
    assert np.array_equal(sa['face'], sb['face']) and np.array_equal(sa['normal'], sb['normal'])
    for rel in ('scenes/basic.json', 'scenes/halfbox_sphere_cube.json'):
        mine = ingest.load_scene(os.path.join(REF, rel))
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            theirs = rrender.load_scene(os.path.join(REF, rel))
        for k in ('face', 'normal', 'material_idx'):
            assert np.allclose(mine['objects']['triangle'][k], theirs['objects']['triangle'][k], rtol=0, atol=1e-15), (rel, k)
        assert mine['camera'] == theirs['camera'] and mine['lights'] == theirs['lights']
        ta = ingest.make_torch_var(mine, device='cpu')
        tb = rrender.make_torch_var(theirs)
        assert torch.equal(ta['objects']['triangle']['face'], tb['objects']['triangle']['face'].cpu())
        assert ta['camera']['viewport'].dtype == tb['camera']['viewport'].dtype
