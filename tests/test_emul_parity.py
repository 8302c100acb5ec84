"""CPU checks of the kernels' math (surf_math.cuh) through the g++-compiled emulation (tests/emul):
conservative filters, exact tests, resolve/shading and the analytic backward, against the reference goldens.
The CUDA kernels evaluate the same inline functions; the GPU parity tests (-m gpu) then pin the kernels."""
import os

import numpy as np
import pytest
import torch

import emul_driver
import parity
import scene_io
from conftest import GOLDEN_DIR, golden_cases


def _load(name):
    return scene_io.load_case(os.path.join(GOLDEN_DIR, name + '.npz'))


@pytest.mark.parametrize('name', golden_cases())
def test_emulated_forward_matches_reference_golden(name):
    scene, params, outs, grads, extra = _load(name)
    res, filter_misses, m = emul_driver.forward(scene, **params)
    assert filter_misses == 0, 'conservative filter rejected %d exact hits' % filter_misses
    ortho = None
    if m.proj == 1:
        from oracle import torch_oracle
        ortho, _, _, _ = torch_oracle.make_rays(scene['camera'])
    ref = {k: outs[k] for k in ('nearest', 'depth', 'image', 'pos', 'normal', 'ray_dir')}
    rep = parity.compare_forward(res, ref, scene, ortho_origins=ortho)
    assert rep['hit_pixels'] == extra['hit_pixels']


@pytest.mark.parametrize('name', [n for n in golden_cases()])
def test_emulated_backward_matches_reference_autograd(name):
    scene, params, outs, grads, extra = _load(name)
    if not grads:
        pytest.skip('forward-only fixture')
    m_res, _, m = emul_driver.forward(scene, **params)
    H, W = outs['depth'].shape
    w = scene_io.loss_weights((H, W), extra['loss_seed'])
    far = scene['camera']['far']
    nearest = torch.tensor(outs['nearest']).reshape(-1)      # forced-winner mode
    depth = torch.tensor(outs['depth']).reshape(-1)
    hit = (depth <= far).float()
    gm = hit[:, None] if extra['hit_only_geom'] else torch.ones_like(hit)[:, None]
    gouts = {'image': w['image'].reshape(-1, 3).contiguous(), 'depth': (w['depth'].reshape(-1) * hit).contiguous(),
             'pos': (w['pos'].reshape(-1, 3) * gm).contiguous(), 'normal': (w['normal'].reshape(-1, 3) * gm).contiguous()}
    g = emul_driver.backward(m, params, nearest, depth, gouts)
    parity.compare_grads(g, grads)


def test_filters_are_conservative_on_dense_random_scenes():
    from surf_renderer_b200 import scenes as synth
    for seed in range(5):
        scene = synth.random_mixed_scene(40 + seed, width=64, height=48, n_disk=60, n_sphere=8, n_tri=40)
        _, filter_misses, _ = emul_driver.forward(scene)
        assert filter_misses == 0
    scene = synth.config_e(m=2000, width=96, height=96, radius=0.02)
    _, filter_misses, _ = emul_driver.forward(scene)
    assert filter_misses == 0
