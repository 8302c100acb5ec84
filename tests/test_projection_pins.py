"""The reference's only executable assertions on render() output - the four consistency tests of
diffrend/torch/projection_layer.py (:428, :461, :609, :846) - run UNMODIFIED against an image the
B200 renderer produced (tests/golden/b200_halfbox_160x120.npz, exported on the GPU box by
tools/projection_consistency.py).  Needs the reference checkout, so it runs in the build container only; the GPU
suite (test_gpu_parity.py::test_stored_b200_render_is_current) ties the stored file to the current kernels."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT

STORED = os.path.join(GOLDEN_DIR, 'b200_halfbox_160x120.npz')


def test_stored_render_is_well_formed():
    z = np.load(STORED)
    assert z['image'].shape == (120, 160, 3) and z['pos'].shape == (120, 160, 3) and z['nearest'].shape == (120, 160)
    assert np.isfinite(z['image']).all() and np.isfinite(z['pos']).all()
    assert len(np.unique(z['nearest'])) > 50


@pytest.mark.skipif(not os.path.isdir('/root/reference/diffrend'), reason='needs the reference checkout')
def test_reference_projection_layer_tests_accept_the_b200_render():
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import projection_consistency
    ran = projection_consistency.check(STORED, verbose=False)
    assert ran == ['test_raster_coordinates', 'test_render_projection_consistency',
                   'test_transformation_consistency', 'test_depth_to_world_consistency']
