"""Pin the CPU oracle (oracle/torch_oracle.py) against golden vectors produced by the real reference
(tests/golden/make_golden.py).  Bit-exact for forward outputs on this CPU build of torch; gradients
within 1e-6 relative (autograd accumulation order is the same graph, but allow last-ulp noise)."""
import os

import numpy as np
import pytest
import torch

import scene_io
from conftest import GOLDEN_DIR, golden_cases
from oracle import torch_oracle


def _run_oracle(name):
    scene, params, outs, grads, extra = scene_io.load_case(os.path.join(GOLDEN_DIR, name + '.npz'))
    sc = scene_io.clone_scene(scene, requires_grad=bool(grads))
    res = torch_oracle.render(sc, **params)
    return sc, params, outs, grads, extra, res


@pytest.mark.parametrize('name', golden_cases())
def test_oracle_forward_matches_reference_golden(name):
    sc, params, outs, grads, extra, res = _run_oracle(name)
    assert np.array_equal(res['nearest'].numpy(), outs['nearest']), 'nearest index differs'
    for k in ('depth', 'ray_dir', 'pos', 'normal', 'image'):
        got = res[k].detach().numpy()
        exp = outs[k]
        same = np.array_equal(got, exp, equal_nan=True)
        if not same:
            # Different host CPUs may select different MKL sgemm kernels (K=3 dot products): allow 2 ulp.
            np.testing.assert_allclose(got, exp, rtol=3e-7, atol=1e-7, equal_nan=True, err_msg=k)


@pytest.mark.parametrize('name', [n for n in golden_cases()])
def test_oracle_gradients_match_reference_golden(name):
    sc, params, outs, grads, extra, res = _run_oracle(name)
    if not grads:
        pytest.skip('forward-only fixture')
    H, W = res['depth'].shape
    w = scene_io.loss_weights((H, W), extra['loss_seed'])
    loss = scene_io.weighted_loss(res, w, sc['camera']['far'], hit_only_geom=extra['hit_only_geom'])
    leaves = scene_io.grad_leaves(sc)
    names = [k for k in grads]
    gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    for k, g in zip(names, gs):
        exp = grads[k]
        got = g.numpy()
        scale = max(1e-12, float(np.nanmax(np.abs(exp))))
        np.testing.assert_allclose(got / scale, exp / scale, rtol=1e-5, atol=1e-6, equal_nan=True, err_msg=k)


@pytest.mark.skipif(not os.path.isdir('/root/reference/diffrend'), reason='reference tree not present')
def test_oracle_matches_live_reference_on_fresh_random_scene():
    """When the reference is importable (this container only), compare on a scene that is NOT a fixture."""
    import sys
    sys.path.insert(0, '/root/reference')
    from diffrend.torch.renderer import render as ref_render
    from surf_renderer_b200 import scenes as synth
    for seed, kw in ((101, {}), (102, {'double_sided': True, 'use_quartic': True, 'tile_size': 333})):
        scene = synth.random_mixed_scene(seed, width=37, height=29, homogeneous=bool(seed % 2))
        a = ref_render(scene_io.clone_scene(scene), **kw)
        b = torch_oracle.render(scene_io.clone_scene(scene), **kw)
        for k in ('nearest', 'depth', 'pos', 'normal', 'image', 'ray_dir'):
            assert torch.equal(a[k], b[k]) or np.array_equal(a[k].numpy(), b[k].numpy(), equal_nan=True), k


from conftest import along_ray_cases   # noqa: E402


@pytest.mark.parametrize('name', along_ray_cases())
def test_oracle_along_ray_matches_reference_golden(name):
    scene, params, outs, grads, extra = scene_io.load_case(os.path.join(GOLDEN_DIR, name + '.npz'))
    sc = scene_io.clone_scene(scene, requires_grad=True)
    if sc['objects']['disk'].get('normal', 1) is None:
        sc['objects']['disk'].pop('normal')
    if 'light_vis' in sc['objects']['disk']:
        sc['objects']['disk']['light_vis'] = sc['objects']['disk']['light_vis'].detach()
    res = torch_oracle.render_along_ray(sc, **params)
    for k in ('image', 'depth', 'pos', 'normal'):
        assert np.array_equal(res[k].detach().numpy(), outs[k], equal_nan=True), k
    H, W = res['depth'].shape
    w = scene_io.loss_weights((H, W), extra['loss_seed'])
    if extra.get('mask_border'):
        m = torch.zeros(H, W)
        m[1:-1, 1:-1] = 1
        w = {k: v * (m[..., None] if v.dim() == 3 else m) for k, v in w.items()}
    loss = sum((res[k] * w[k]).sum() for k in ('image', 'depth', 'pos', 'normal'))
    leaves = {'objects/disk/pos': sc['objects']['disk']['pos'], 'objects/disk/normal': sc['objects']['disk'].get('normal'),
              'materials/albedo': sc['materials']['albedo'], 'materials/coeffs': sc['materials']['coeffs'],
              'lights/pos': sc['lights']['pos'], 'lights/attenuation': sc['lights']['attenuation'],
              'lights/ambient': sc['lights']['ambient'], 'colors': sc['colors']}
    gs = torch.autograd.grad(loss, [leaves[k] for k in grads], allow_unused=True)
    for k, g in zip(grads, gs):
        scale = max(1e-12, float(np.abs(grads[k]).max()))
        np.testing.assert_allclose(g.numpy() / scale, grads[k] / scale, rtol=1e-5, atol=1e-6, err_msg=k)


@pytest.mark.skipif(not os.path.isdir('/root/reference/diffrend'), reason='reference tree not present')
def test_oracle_norm_depth_image_only_matches_live_reference():
    """renderer.py:245-260 runs in the reference with tiled=False on triangle scenes (their intersector ignores
    disable_normals)."""
    import sys
    sys.path.insert(0, '/root/reference')
    from diffrend.torch.renderer import render as ref_render
    scene, params, outs, grads, extra = scene_io.load_case(os.path.join(GOLDEN_DIR, 'c_torus_64.npz'))
    a = ref_render(scene_io.clone_scene(scene), tiled=False, norm_depth_image_only=True, double_sided=True)
    b = torch_oracle.render(scene_io.clone_scene(scene), tiled=False, norm_depth_image_only=True, double_sided=True)
    assert torch.equal(a['image'], b['image']) and torch.equal(a['depth'], b['depth']) and torch.equal(a['nearest'], b['nearest'])


@pytest.mark.skipif(not os.path.isdir('/root/reference/diffrend'), reason='reference tree not present')
def test_oracle_along_ray_norm_depth_matches_live_reference():
    """renderer.py:677-686, the depth-only branch of render_splats_along_ray."""
    import sys
    sys.path.insert(0, '/root/reference')
    sys.path.insert(0, GOLDEN_DIR)
    from diffrend.torch.renderer import render_splats_along_ray as ref_fn
    from make_golden_along_ray_scene import along_ray_scene
    scene = along_ray_scene(13, 26, 18)
    scene['camera']['far'] = 4.0                     # some fragments beyond `far`
    a = ref_fn(scene_io.clone_scene(scene), norm_depth_image_only=True)
    b = torch_oracle.render_along_ray(scene_io.clone_scene(scene), norm_depth_image_only=True)
    assert float((a['depth'] >= 4.0).float().mean()) > 0.02
    for k in ('image', 'depth', 'pos', 'normal'):
        assert torch.equal(a[k], b[k]), k
