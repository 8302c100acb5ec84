"""CPU checks (emulation of the kernel math) for the parts the reference cannot pin directly:
  * sphere gradients - NaN in the reference; checked against float64 autograd of the oracle's
    NaN-free sphere variant (oracle-only switch SAFE_SPHERE),
  * shadow rays      - the reference's shadow branch needs CUDA (renderer.py:311); checked against the oracle's
    device-agnostic restatement (itself bit-exact on the reference-generated sh_* shadow fixtures)."""
import numpy as np
import pytest
import torch

import emul_driver
import parity
import scene_io
from oracle import torch_oracle
from surf_renderer_b200 import scenes as synth


def _to64(scene):
    def rec(v):
        if isinstance(v, dict):
            return {k: rec(x) for k, x in v.items()}
        if isinstance(v, torch.Tensor) and v.is_floating_point():
            return v.double()
        return v
    return rec(scene)


@pytest.fixture
def safe_sphere_f64():
    torch_oracle.SAFE_SPHERE = True
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(torch.float32)
    torch_oracle.SAFE_SPHERE = False


def sphere_scene():
    s = synth.random_mixed_scene(21, width=40, height=32, n_disk=3, n_plane=1, n_sphere=5, n_tri=2,
                                 order=('sphere', 'disk', 'triangle', 'plane'))
    return s


def test_sphere_gradients_match_float64_autograd(safe_sphere_f64):
    scene = sphere_scene()
    sc64 = scene_io.clone_scene(_to64(scene), requires_grad=True)
    ref = torch_oracle.render(sc64)
    H, W = ref['depth'].shape
    w = scene_io.loss_weights((H, W), 77)
    w = {k: v.double() for k, v in w.items()}
    far = scene['camera']['far']
    loss = scene_io.weighted_loss(ref, w, far)
    leaves = scene_io.grad_leaves(sc64)
    names = [k for k, v in leaves.items() if v.requires_grad]
    gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    ref_g = {k: g.detach().numpy() for k, g in zip(names, gs) if g is not None}
    assert np.isfinite(ref_g['objects/sphere/pos']).all() and np.abs(ref_g['objects/sphere/radius']).max() > 0
    torch.set_default_dtype(torch.float32)
    res, misses, m = emul_driver.forward(scene)
    assert misses == 0
    # same winners as the float64 oracle (no ties in this scene) -> forced-winner backward
    assert np.array_equal(res['nearest'].numpy(), ref['nearest'].numpy())
    nearest = ref['nearest'].reshape(-1)
    depth = ref['depth'].reshape(-1).float()
    hit = (depth <= far).float()
    gouts = {'image': w['image'].float().reshape(-1, 3).contiguous(), 'depth': (w['depth'].float().reshape(-1) * hit).contiguous(),
             'pos': (w['pos'].float().reshape(-1, 3) * hit[:, None]).contiguous(),
             'normal': (w['normal'].float().reshape(-1, 3) * hit[:, None]).contiguous()}
    g = emul_driver.backward(m, {}, nearest, depth, gouts)
    # fp32 kernel math vs a float64 reference: silhouette pixels (G.d -> 0) amplify fp32 rounding, so the sphere
    # leaves get 2e-4 of the leaf's max |grad| as absolute slack; everything else keeps the normal bound
    parity.compare_grads(g, ref_g, rtol=2e-4, atol_scale=2e-5, skip=('objects/sphere/pos', 'objects/sphere/radius'))
    sph = {k: ref_g[k] for k in ('objects/sphere/pos', 'objects/sphere/radius')}
    worst = parity.compare_grads(g, sph, rtol=1e-3, atol_scale=2e-4)
    assert worst['objects/sphere/radius'] < 2e-4


def test_sphere_forward_matches_reference_semantics():
    """fp32 forward with spheres agrees with the stock oracle (reference semantics) on this scene."""
    scene = sphere_scene()
    res, misses, _ = emul_driver.forward(scene)
    ref = torch_oracle.render(scene_io.clone_scene(scene))
    parity.compare_forward(res, {k: v for k, v in ref.items() if isinstance(v, torch.Tensor)}, scene)


@pytest.mark.parametrize('seed', [31, 32])
def test_shadow_visibility_matches_oracle_restatement(seed):
    scene = synth.random_mixed_scene(seed, width=36, height=28, n_disk=10, n_sphere=0, n_tri=6, n_plane=1)
    res, misses, m = emul_driver.forward(scene, shadow=True)
    assert misses == 0, 'a conservative filter (camera or shadow rays) rejected %d exact hits' % misses
    ref = torch_oracle.render(scene_io.clone_scene(scene), shadow=True)
    rep = parity.compare_forward(res, {k: v for k, v in ref.items() if isinstance(v, torch.Tensor)}, scene, atol=2e-5)
    no_shadow, _, _ = emul_driver.forward(scene)
    assert not torch.equal(no_shadow['image'], res['image']), 'scene casts no shadows - test is vacuous'
    # gradients with shadows: visibility is a constant mask
    sc = scene_io.clone_scene(scene, requires_grad=True)
    r2 = torch_oracle.render(sc, shadow=True)
    H, W = r2['depth'].shape
    w = scene_io.loss_weights((H, W), seed)
    far = scene['camera']['far']
    kink = parity.kink_mask(scene, {k: v.detach() for k, v in r2.items() if isinstance(v, torch.Tensor)}, {})
    good = torch.tensor(rep['good_mask'] & ~kink).view(H, W)
    for k in w:
        w[k] = w[k] * (good[..., None] if w[k].dim() == 3 else good)
    loss = scene_io.weighted_loss(r2, w, far)
    leaves = scene_io.grad_leaves(sc)
    names = [k for k, v in leaves.items() if v.requires_grad]
    gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    ref_g = {k: g.numpy() for k, g in zip(names, gs) if g is not None}
    hit = (r2['depth'].detach().reshape(-1) <= far).float()
    gouts = {'image': w['image'].reshape(-1, 3).contiguous(), 'depth': (w['depth'].reshape(-1) * hit).contiguous(),
             'pos': (w['pos'].reshape(-1, 3) * hit[:, None]).contiguous(),
             'normal': (w['normal'].reshape(-1, 3) * hit[:, None]).contiguous()}
    g = emul_driver.backward(m, {'shadow': True}, r2['nearest'].detach().reshape(-1), r2['depth'].detach().reshape(-1), gouts)
    parity.compare_grads(g, ref_g)


def _random_case(seed):
    """shared with tests/test_gpu_parity.py::test_randomized_scenes_and_options_vs_oracle"""
    from surf_renderer_b200 import scenes as synth
    rng = np.random.RandomState(seed)
    with_grads = seed % 2 == 0
    orders = [('disk', 'sphere', 'triangle', 'plane'), ('triangle', 'plane', 'disk', 'sphere'), ('plane', 'disk', 'triangle', 'sphere')]
    ortho = (not with_grads) and rng.rand() < 0.3
    width, height = int(rng.randint(17, 90)), int(rng.randint(9, 70))
    if ortho:
        width, height = min(width, 60), min(height, 45)                 # the reference's ortho path: one tile of pixels
    scene = synth.random_mixed_scene(seed, width=width, height=height, n_disk=int(rng.randint(5, 60)), n_tri=int(rng.randint(3, 40)),
                                     n_sphere=0 if with_grads else int(rng.randint(1, 6)), n_plane=int(rng.randint(0, 3)),
                                     n_lights=int(rng.randint(1, 6)), order=orders[seed % 3], homogeneous=bool(rng.rand() < 0.5),
                                     proj='orthographic' if ortho else 'perspective')
    params = {'double_sided': bool(rng.rand() < 0.5), 'use_quartic': bool(rng.rand() < 0.3)}
    if not ortho and rng.rand() < 0.4:
        params['shadow'] = True
    mode = int(rng.choice([0, 0, 2, 3, 4])) if not ortho else 0
    return scene, params, mode, with_grads, ortho


@pytest.mark.parametrize('seed', list(range(200, 230)))
def test_emulated_randomized_scenes_and_options_vs_oracle(seed):
    """The option-space sweep of the GPU suite on the CPU emulation of the kernel math: random scenes with every
    primitive kind, random viewport / projection / primitive order / layouts / kwargs, forward against the oracle;
    the conservative filters never reject an exact hit."""
    import emul_driver
    import parity
    import scene_io
    from oracle import torch_oracle
    scene, params, _mode, _with_grads, ortho = _random_case(seed)
    res, filter_misses, m = emul_driver.forward(scene, **params)
    assert filter_misses == 0
    ref = torch_oracle.render(scene_io.clone_scene(scene), **params)
    origins = torch_oracle.make_rays(scene['camera'])[0] if ortho else None
    ref_np = {k: v.detach() for k, v in ref.items() if isinstance(v, torch.Tensor)}
    parity.compare_forward(res, ref_np, scene, ortho_origins=origins)
