"""oracle/numpy_oracle.py (the restated numpy twin renderer, the second CPU baseline of bench.py) against outputs of
the reference's diffrend/numpy/renderer.py::render stored in tests/golden/np_*.npz."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN_DIR, numpy_twin_cases
from oracle import numpy_oracle, torch_oracle

sys.path.insert(0, GOLDEN_DIR)
from make_golden_numpy import load_np_case   # noqa: E402


def test_fixture_list_is_complete():
    assert len(numpy_twin_cases()) >= 5


@pytest.mark.parametrize('name', numpy_twin_cases())
def test_numpy_oracle_reproduces_reference_outputs(name):
    scene, outs = load_np_case(os.path.join(GOLDEN_DIR, name + '.npz'))
    res = numpy_oracle.render(scene)
    assert np.array_equal(res['nearest'], outs['nearest'])
    # float64 throughout and the same operation order: equal to the last bit up to BLAS summation order of the K=4
    # dot products (np.dot may pick another kernel on another CPU), hence 1e-12 instead of array_equal
    for k in ('depth', 'image', 'ray_dir'):
        a, b = res[k], outs[k]
        assert a.shape == b.shape and a.dtype == b.dtype == np.float64
        assert np.array_equal(np.isfinite(a), np.isfinite(b)), k
        f = np.isfinite(b)
        assert np.allclose(a[f], b[f], rtol=1e-12, atol=1e-13), k


def test_pixel_subset_matches_full_frame():
    scene, outs = load_np_case(os.path.join(GOLDEN_DIR, 'np_random_mixed_48x36.npz'))
    idx = np.arange(5, 48 * 36, 7)
    res = numpy_oracle.render(scene, pixel_subset=idx)
    assert np.array_equal(res['nearest'], outs['nearest'].reshape(-1)[idx])
    f = np.isfinite(outs['depth'].reshape(-1)[idx])
    assert np.allclose(res['depth'][f], outs['depth'].reshape(-1)[idx][f], rtol=1e-12)
    assert np.allclose(res['image'], outs['image'].reshape(-1, 3)[idx], rtol=1e-12, atol=1e-13)


def test_numpy_twin_and_torch_renderer_agree_on_visibility():
    """Where the two camera conventions coincide (up perpendicular to the view direction, ops.py:105-111 vs
    torch/utils.py:402-427) both reference renderers see the same splat at (nearly) every pixel - a cross-check of
    the two restatements against each other, float64 vs float32."""
    from surf_renderer_b200 import scenes as synth
    scene = synth.config_e(m=2000, width=40, height=40, radius=0.03)
    a = numpy_oracle.render(numpy_oracle.homogeneous_scene(scene))
    b = torch_oracle.render(scene)
    hit_a, hit_b = np.isfinite(a['depth']), (b['depth'].numpy() <= scene['camera']['far'])
    assert (hit_a != hit_b).mean() < 0.01
    both = hit_a & hit_b
    same = a['nearest'][both] == b['nearest'].numpy()[both]
    assert same.mean() > 0.99
    assert np.allclose(a['depth'][both][same], b['depth'].numpy()[both][same], rtol=1e-5)
