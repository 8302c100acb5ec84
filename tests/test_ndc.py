"""render_splats_NDC (reference renderer.py:358-474): the oracle restatement pinned on reference-generated fixtures
(CPU), and the CUDA path through the C ABI against the fixtures and the oracle (GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

import parity
import scene_io
from conftest import GOLDEN_DIR, ndc_cases

sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
from make_golden_ndc_scene import ndc_scene       # noqa: E402

LEAVES = ('objects/disk/pos', 'objects/disk/normal', 'materials/albedo', 'materials/coeffs', 'lights/pos',
          'lights/attenuation', 'lights/ambient', 'colors')


def _load(name):
    return scene_io.load_case(os.path.join(GOLDEN_DIR, name + '.npz'))


def _leaf(sc, key):
    d = sc
    for p in key.split('/'):
        d = d[p]
    return d


def _loss(res, H, W, seed, dev='cpu'):
    w = scene_io.loss_weights((H, W), seed)
    return sum((res[k] * w[k].to(dev)).sum() for k in ('image', 'depth', 'pos', 'normal'))


@pytest.mark.parametrize('name', ndc_cases())
def test_oracle_ndc_is_bit_exact_on_reference_golden(name):
    from oracle import torch_oracle
    scene, params, outs, grads, extra = _load(name)
    sc = scene_io.clone_scene(scene, requires_grad=True)
    res = torch_oracle.render_splats_ndc(sc, **params)
    for k in ('image', 'depth', 'pos', 'normal'):
        assert np.array_equal(res[k].detach().numpy(), outs[k]), k
    H, W = res['depth'].shape
    gs = torch.autograd.grad(_loss(res, H, W, extra['loss_seed']), [_leaf(sc, k) for k in grads])
    for k, g in zip(grads, gs):
        assert np.array_equal(g.numpy(), grads[k]), k


@pytest.mark.skipif(not os.path.isdir('/root/reference/diffrend'), reason='reference tree not present')
def test_oracle_ndc_matches_live_reference_norm_depth():
    sys.path.insert(0, '/root/reference')
    from diffrend.torch.renderer import render_splats_NDC as ref_fn
    from oracle import torch_oracle
    scene = ndc_scene(7, 12, 9)
    scene['camera']['far'] = 5.0
    a = ref_fn(scene_io.clone_scene(scene), norm_depth_image_only=True)
    b = torch_oracle.render_splats_ndc(scene_io.clone_scene(scene), norm_depth_image_only=True)
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_ndc_argument_checks_need_no_gpu():
    """host-side validation runs before any device work"""
    from surf_renderer_b200.along_ray import _NDCCall
    scene = ndc_scene(3, 8, 6)
    call = _NDCCall(scene, {}, torch.device('cpu'))
    assert call.ndc_stride == 3 and call.n_src == 48 and call.n_lights == scene['lights']['pos'].shape[0]
    bad = scene_io.clone_scene(scene)
    bad['objects']['disk']['pos'] = bad['objects']['disk']['pos'][:-1]
    with pytest.raises(RuntimeError):
        _NDCCall(bad, {}, torch.device('cpu'))
    bad = scene_io.clone_scene(scene)
    bad['lights']['pos'] = bad['lights']['pos'][:, :3]
    with pytest.raises(ValueError):
        _NDCCall(bad, {}, torch.device('cpu'))


@pytest.mark.gpu
@pytest.mark.parametrize('name', ndc_cases())
def test_gpu_ndc_matches_reference_golden(name):
    import surf_renderer_b200
    scene, params, outs, grads, extra = _load(name)
    sc = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
    res = surf_renderer_b200.render_splats_NDC(sc, **params)
    for k in ('image', 'depth', 'pos', 'normal'):
        a, b = res[k].detach().cpu().numpy().astype(np.float64), outs[k].astype(np.float64)
        assert a.shape == b.shape
        err = np.abs(a - b) - (parity.ATOL + parity.RTOL * np.abs(b))
        assert not (err > 0).any(), '%s: max abs diff %.3g' % (k, np.abs(a - b).max())
    H, W = res['depth'].shape
    gs = torch.autograd.grad(_loss(res, H, W, extra['loss_seed'], 'cuda'), [_leaf(sc, k) for k in grads])
    parity.compare_grads({k: g.cpu() for k, g in zip(grads, gs)}, grads)


@pytest.mark.gpu
@pytest.mark.parametrize('seed', list(range(500, 506)))
def test_gpu_ndc_randomized_vs_oracle(seed):
    import surf_renderer_b200
    from oracle import torch_oracle
    rng = np.random.RandomState(seed)
    W, H = int(rng.randint(8, 70)), int(rng.randint(6, 50))
    scene = ndc_scene(seed, W, H, homogeneous=bool(rng.rand() < 0.5), mats=int(rng.randint(1, 5)))
    params = {'double_sided': bool(rng.rand() < 0.5), 'use_quartic': bool(rng.rand() < 0.3)}
    sc = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
    osc = scene_io.clone_scene(scene, requires_grad=True)
    res = surf_renderer_b200.render_splats_NDC(sc, **params)
    ref = torch_oracle.render_splats_ndc(osc, **params)
    for k in ('image', 'depth', 'pos', 'normal'):
        assert torch.allclose(res[k].cpu().double(), ref[k].detach().double(), rtol=parity.RTOL, atol=parity.ATOL), k
    gs = torch.autograd.grad(_loss(res, H, W, seed, 'cuda'), [_leaf(sc, k) for k in LEAVES])
    gr = torch.autograd.grad(_loss(ref, H, W, seed), [_leaf(osc, k) for k in LEAVES])
    parity.compare_grads({k: g.cpu() for k, g in zip(LEAVES, gs)}, {k: g for k, g in zip(LEAVES, gr)})


@pytest.mark.gpu
def test_gpu_ndc_norm_depth_image_only():
    """renderer.py:392-401: normalised depth image, homogeneous [N,4] positions and the caller's normals returned"""
    import surf_renderer_b200
    from oracle import torch_oracle
    scene = ndc_scene(11, 21, 17)
    scene['camera']['far'] = 5.0
    ref = torch_oracle.render_splats_ndc(scene_io.clone_scene(scene), norm_depth_image_only=True)
    res = surf_renderer_b200.render_splats_NDC(scene_io.clone_scene(scene, device='cuda'), norm_depth_image_only=True)
    assert res['image'].shape == (17, 21) and res['pos'].shape == (21 * 17, 4)
    for k in ('image', 'depth', 'pos', 'normal'):
        assert torch.allclose(res[k].cpu().double(), ref[k].double(), rtol=parity.RTOL, atol=parity.ATOL), k
