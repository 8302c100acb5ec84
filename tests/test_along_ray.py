"""render_splats_along_ray (reference renderer.py:537-751): CPU checks of the kernel math through the
emulation against reference-generated goldens, and GPU parity tests through the C ABI."""
import os

import numpy as np
import pytest
import torch

import emul_driver
import parity
import scene_io
from conftest import GOLDEN_DIR, along_ray_cases


def _load(name):
    return scene_io.load_case(os.path.join(GOLDEN_DIR, name + '.npz'))


def _check_outputs(res, outs):
    for k in ('image', 'depth', 'pos', 'normal'):
        a, b = res[k].detach().cpu().numpy().astype(np.float64), outs[k].astype(np.float64)
        err = np.abs(a - b) - (parity.ATOL + parity.RTOL * np.abs(b))
        assert not (err > 0).any(), '%s: max abs diff %.3g' % (k, np.abs(a - b).max())


def _gouts(H, W, seed):
    w = scene_io.loss_weights((H, W), seed)
    return {k: w[k].reshape(H * W, -1).squeeze(-1).contiguous() if k == 'depth' else w[k].reshape(H * W, 3).contiguous()
            for k in ('image', 'depth', 'pos', 'normal')}


@pytest.mark.parametrize('name', along_ray_cases())
def test_emulated_along_ray_matches_reference_golden(name):
    scene, params, outs, grads, extra = _load(name)
    res, inp = emul_driver.splats_forward(scene, **params)
    _check_outputs(res, outs)
    g = emul_driver.splats_backward(scene, params, _gouts(inp.height, inp.width, extra['loss_seed']))
    parity.compare_grads(g, grads)


def test_along_ray_unsupported_options_raise():
    import surf_renderer_b200
    scene, params, outs, grads, extra = _load(along_ray_cases()[0])
    with pytest.raises(NotImplementedError):
        surf_renderer_b200.render_splats_along_ray(scene, samples=2)
    sc = scene_io.clone_scene(scene)
    sc['objects']['disk']['normal'] = None
    with pytest.raises((NotImplementedError, RuntimeError)):
        surf_renderer_b200.render_splats_along_ray(sc)


@pytest.mark.gpu
@pytest.mark.parametrize('name', along_ray_cases())
def test_gpu_along_ray_matches_reference_golden(name):
    import surf_renderer_b200
    scene, params, outs, grads, extra = _load(name)
    sc = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
    if 'light_vis' in sc['objects']['disk']:
        sc['objects']['disk']['light_vis'] = sc['objects']['disk']['light_vis'].detach()
    res = surf_renderer_b200.render_splats_along_ray(sc, **params)
    _check_outputs(res, outs)
    H, W = res['depth'].shape
    w = scene_io.loss_weights((H, W), extra['loss_seed'])
    loss = sum((res[k] * w[k].cuda()).sum() for k in ('image', 'depth', 'pos', 'normal'))
    leaves = {'objects/disk/pos': sc['objects']['disk']['pos'], 'objects/disk/normal': sc['objects']['disk']['normal'],
              'materials/albedo': sc['materials']['albedo'], 'materials/coeffs': sc['materials']['coeffs'],
              'lights/pos': sc['lights']['pos'], 'lights/attenuation': sc['lights']['attenuation'],
              'lights/ambient': sc['lights']['ambient'], 'colors': sc['colors']}
    gs = torch.autograd.grad(loss, [leaves[k] for k in grads], allow_unused=True)
    parity.compare_grads({k: g.cpu() for k, g in zip(grads, gs)}, grads)


@pytest.mark.gpu
def test_gpu_along_ray_gan_shape_vs_oracle():
    """GAN generator shape: 128x128 splats, 1 material, against the oracle restatement; fwd + bwd."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    from make_golden_along_ray_scene import along_ray_scene
    import surf_renderer_b200
    from oracle import torch_oracle
    scene = along_ray_scene(71, 128, 128, mats=1)
    sc = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
    res = surf_renderer_b200.render_splats_along_ray(sc)
    osc = scene_io.clone_scene(scene, requires_grad=True)
    ref = torch_oracle.render_along_ray(osc)
    _check_outputs(res, {k: v.detach().numpy() for k, v in ref.items()})
    (res['image'].sum() + res['depth'].sum()).backward()
    (ref['image'].sum() + ref['depth'].sum()).backward()
    parity.compare_grads({'z': sc['objects']['disk']['pos'].grad.cpu(), 'n': sc['objects']['disk']['normal'].grad.cpu(),
                          'l': sc['lights']['pos'].grad.cpu()},
                         {'z': osc['objects']['disk']['pos'].grad, 'n': osc['objects']['disk']['normal'].grad,
                          'l': osc['lights']['pos'].grad})
