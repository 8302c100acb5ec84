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
    # 'avg_normal' estimates are rounding noise where the eight cross products cancel (image-border pixels under the
    # reflect padding): the reference's own normal has length << 1 there.  Those pixels are excluded.
    ok = np.linalg.norm(outs['normal'].astype(np.float64), axis=-1) > 0.5
    assert ok.mean() > 0.8
    for k in ('image', 'depth', 'pos', 'normal'):
        a, b = res[k].detach().cpu().numpy().astype(np.float64), outs[k].astype(np.float64)
        err = np.abs(a - b) - (parity.ATOL + parity.RTOL * np.abs(b))
        err = err[ok]
        assert not (err > 0).any(), '%s: max abs diff %.3g' % (k, np.abs(a - b)[ok].max())


def _weights(H, W, extra):
    w = scene_io.loss_weights((H, W), extra['loss_seed'])
    if extra.get('mask_border'):
        m = torch.zeros(H, W)
        m[1:-1, 1:-1] = 1
        w = {k: v * (m[..., None] if v.dim() == 3 else m) for k, v in w.items()}
    return w


def _gouts(H, W, extra):
    w = _weights(H, W, extra)
    return {k: w[k].reshape(H * W, -1).squeeze(-1).contiguous() if k == 'depth' else w[k].reshape(H * W, 3).contiguous()
            for k in ('image', 'depth', 'pos', 'normal')}


def _leaf_map(sc):
    return {'objects/disk/pos': sc['objects']['disk']['pos'], 'objects/disk/normal': sc['objects']['disk'].get('normal'),
            'materials/albedo': sc['materials']['albedo'], 'materials/coeffs': sc['materials']['coeffs'],
            'lights/pos': sc['lights']['pos'], 'lights/attenuation': sc['lights']['attenuation'],
            'lights/ambient': sc['lights']['ambient'], 'colors': sc['colors']}


def _prepared(scene, requires_grad=False, device='cpu'):
    sc = scene_io.clone_scene(scene, device=device, requires_grad=requires_grad)
    if sc['objects']['disk'].get('normal', 1) is None:
        sc['objects']['disk'].pop('normal')
    if 'light_vis' in sc['objects']['disk']:
        sc['objects']['disk']['light_vis'] = sc['objects']['disk']['light_vis'].detach()
    return sc


@pytest.mark.parametrize('name', along_ray_cases())
def test_emulated_along_ray_matches_reference_golden(name):
    """kernel math through the emulation; for estimated normals / supersampling the fragments come from the package's
    tensor program (run on CPU tensors here) and the emulated kernel gradients are chained through it with autograd."""
    import along_ray_program as along_ray
    scene, params, outs, grads, extra = _load(name)
    sc = _prepared(scene, requires_grad=True)
    inp = along_ray.build_inputs(sc, params, torch.device('cpu'))
    frag_pos, frag_n = inp.floats[0], inp.floats[1]
    if inp.explicit:
        inp.floats = [t.detach().contiguous() for t in inp.floats]
    res, _ = emul_driver.splats_forward(sc, _inp=inp, **params)
    _check_outputs(res, outs)
    g = emul_driver.splats_backward(sc, params, _gouts(inp.height, inp.width, extra), _inp=inp)
    if inp.explicit:        # chain d/d(fragment pos, normal) back to z (and the given normals) through the tensor program
        leaves = _leaf_map(sc)
        wanted = [k for k in ('objects/disk/pos', 'objects/disk/normal') if k in grads]
        roots, seeds = [frag_pos], [g['objects/disk/pos']]
        if frag_n.requires_grad:
            roots.append(frag_n); seeds.append(g['objects/disk/normal'])
        gz = torch.autograd.grad(roots, [leaves[k] for k in wanted], seeds, allow_unused=True)
        for k, v in zip(wanted, gz):
            g[k] = v if v is not None else torch.zeros_like(leaves[k])
    parity.compare_grads(g, grads)


@pytest.mark.gpu
def test_gpu_along_ray_norm_depth_image_only():
    """renderer.py:677-686: normalised depth image; fragments at or beyond `far` take the minimum."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    from make_golden_along_ray_scene import along_ray_scene
    import surf_renderer_b200
    from oracle import torch_oracle
    scene = along_ray_scene(13, 26, 18)
    scene['camera']['far'] = 4.0
    ref = torch_oracle.render_along_ray(scene_io.clone_scene(scene), norm_depth_image_only=True)
    res = surf_renderer_b200.render_splats_along_ray(scene_io.clone_scene(scene, device='cuda'), norm_depth_image_only=True)
    assert res['image'].shape == (18, 26) and res['pos'].shape == (26 * 18, 3)
    for k in ('image', 'depth', 'pos', 'normal'):
        a, b = res[k].cpu().double(), ref[k][..., :3].double() if k in ('pos', 'normal') else ref[k].double()
        assert torch.allclose(a, b, rtol=parity.RTOL, atol=parity.ATOL), k


@pytest.mark.gpu
@pytest.mark.parametrize('name', along_ray_cases())
def test_gpu_along_ray_matches_reference_golden(name):
    import surf_renderer_b200
    scene, params, outs, grads, extra = _load(name)
    sc = _prepared(scene, requires_grad=True, device='cuda')
    res = surf_renderer_b200.render_splats_along_ray(sc, **params)
    _check_outputs(res, outs)
    H, W = res['depth'].shape
    w = _weights(H, W, extra)
    loss = sum((res[k] * w[k].cuda()).sum() for k in ('image', 'depth', 'pos', 'normal'))
    leaves = _leaf_map(sc)
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


@pytest.mark.gpu
@pytest.mark.parametrize('seed', list(range(300, 310)))
def test_gpu_along_ray_randomized_options_vs_oracle(seed):
    """Option-space sweep of render_splats_along_ray against the oracle restatement (itself bit-exact on the six
    reference-generated fixtures): random viewport, z as [N] or [N,3], per-light visibility, materials, use_quartic,
    given normals or estimated ones ('plane' / 'avg_normal' on a smooth depth field), samples 1 or 2; outputs and the
    gradients of depths, normals and light positions."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    from make_golden_along_ray_scene import along_ray_scene
    import surf_renderer_b200
    from oracle import torch_oracle
    rng = np.random.RandomState(seed)
    W, H = int(rng.randint(9, 48)), int(rng.randint(7, 40))
    estimate = rng.rand() < 0.4
    samples = 2 if rng.rand() < 0.3 else 1
    scene = along_ray_scene(seed, W, H, with_vis=bool(rng.rand() < 0.4), pos3=bool(rng.rand() < 0.5),
                            mats=int(rng.randint(1, 5)), smooth_z=estimate or samples > 1)
    params = {'use_quartic': bool(rng.rand() < 0.3)}
    if samples > 1:
        params['samples'] = samples
    if estimate:
        del scene['objects']['disk']['normal']
        params['normal_estimation_method'] = 'plane' if rng.rand() < 0.6 else 'avg_normal'
    sc = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
    osc = scene_io.clone_scene(scene, requires_grad=True)
    res = surf_renderer_b200.render_splats_along_ray(sc, **params)
    ref = torch_oracle.render_along_ray(osc, **params)
    _check_outputs(res, {k: v.detach().numpy() for k, v in ref.items()})
    ok = torch.tensor(np.linalg.norm(ref['normal'].detach().numpy().astype(np.float64), axis=-1) > 0.5)   # see _check_outputs
    w = torch.rand(ref['image'].shape, generator=torch.Generator().manual_seed(seed)) * ok[..., None]
    ((res['image'] * w.cuda()).sum() + (res['depth'] * ok.cuda()).sum()).backward()
    ((ref['image'] * w).sum() + (ref['depth'] * ok).sum()).backward()
    cand = {'z': sc['objects']['disk']['pos'].grad.cpu(), 'l': sc['lights']['pos'].grad.cpu()}
    want = {'z': osc['objects']['disk']['pos'].grad, 'l': osc['lights']['pos'].grad}
    if not estimate:
        cand['n'], want['n'] = sc['objects']['disk']['normal'].grad.cpu(), osc['objects']['disk']['normal'].grad
    parity.compare_grads(cand, want, rtol=1e-4, atol_scale=5e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('variant', ['given_normals', 'plane_samples2', 'avg_normal'])
def test_gpu_along_ray_batch_equals_the_per_element_loop(variant):
    """render_splats_along_ray_batch (one call per direction) against the loop of gan.py:563-597: per element its own
    depths, normals, camera eye and light positions; outputs and gradients - also of the shared material table, which
    receives the sum over the batch."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    from make_golden_along_ray_scene import along_ray_scene
    import surf_renderer_b200
    B, W, H = 5, 24, 20
    g = torch.Generator().manual_seed(9)
    elems = [along_ray_scene(400 + b, W, H, mats=2, smooth_z=True) for b in range(B)]
    params = {}
    if variant == 'plane_samples2':
        params = {'samples': 2, 'normal_estimation_method': 'plane'}
    elif variant == 'avg_normal':
        params = {'normal_estimation_method': 'avg_normal'}
    estimate = variant != 'given_normals'
    eyes = torch.stack([torch.tensor([0.3 * b, 0.1 * b, 4.0 + 0.2 * b, 1.0]) for b in range(B)])
    lights = torch.stack([elems[b]['lights']['pos'] + 0.1 * b for b in range(B)])
    zs = torch.stack([elems[b]['objects']['disk']['pos'] for b in range(B)])
    ns = torch.stack([elems[b]['objects']['disk']['normal'] for b in range(B)])

    def scene_of(z, n, eye, lp, albedo):
        sc = scene_io.clone_scene(elems[0], device='cuda')
        sc['objects']['disk']['pos'] = z
        if estimate:
            sc['objects']['disk'].pop('normal', None)
        else:
            sc['objects']['disk']['normal'] = n
        sc['camera']['eye'] = eye
        sc['lights']['pos'] = lp
        sc['materials']['albedo'] = albedo
        return sc

    K = 2 if 'samples' in params else 1
    w = torch.rand(B, H * K, W * K, 3, generator=g)
    if variant == 'avg_normal':      # border normals are rounding noise there (see _check_outputs): ill-conditioned gradients
        w[:, :2 * K] = 0; w[:, -2 * K:] = 0; w[:, :, :2 * K] = 0; w[:, :, -2 * K:] = 0
    w = w.cuda()
    dmask = (w[..., 0] > 0).float() if variant == 'avg_normal' else torch.ones_like(w[..., 0])
    # batched call
    zb, nb = zs.cuda().requires_grad_(True), ns.cuda().requires_grad_(True)
    lb = lights.cuda().requires_grad_(True)
    alb_b = elems[0]['materials']['albedo'].cuda().requires_grad_(True)
    res = surf_renderer_b200.render_splats_along_ray_batch(scene_of(zb, nb, eyes.cuda(), lb, alb_b), **params)
    ((res['image'] * w).sum() + (res['depth'] * dmask).sum()).backward()
    # loop
    zl, nl = zs.cuda().requires_grad_(True), ns.cuda().requires_grad_(True)
    ll = lights.cuda().requires_grad_(True)
    alb_l = elems[0]['materials']['albedo'].cuda().requires_grad_(True)
    loss = 0
    outs = []
    for b in range(B):
        r = surf_renderer_b200.render_splats_along_ray(scene_of(zl[b], nl[b], eyes[b].cuda(), ll[b], alb_l), **params)
        outs.append(r)
        loss = loss + (r['image'] * w[b]).sum() + (r['depth'] * dmask[b]).sum()
    loss.backward()
    torch.cuda.synchronize()
    for k in ('image', 'depth', 'pos', 'normal'):
        assert torch.equal(res[k], torch.stack([o[k] for o in outs])), k
    assert torch.allclose(zb.grad, zl.grad, rtol=1e-4, atol=1e-5 * float(zl.grad.abs().max()))
    assert torch.allclose(lb.grad, ll.grad, rtol=1e-4, atol=1e-6 * float(ll.grad.abs().max()))
    assert torch.allclose(alb_b.grad, alb_l.grad, rtol=1e-4, atol=1e-6 * float(alb_l.grad.abs().max()))
    if not estimate:
        assert torch.allclose(nb.grad, nl.grad, rtol=1e-4, atol=1e-6 * float(nl.grad.abs().max()))


@pytest.mark.gpu
def test_gpu_along_ray_is_kernels_only():
    """no torch op sits between the caller's z and the kernels: forward + backward of an estimated-normal,
    supersampled frame launch only the library's kernels (counted by the library) and exactly the autograd glue's
    allocations / fills"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    from make_golden_along_ray_scene import along_ray_scene
    import surf_renderer_b200
    from surf_renderer_b200._lib import lib
    scene = along_ray_scene(5, 32, 32, mats=1, smooth_z=True)
    del scene['objects']['disk']['normal']
    sc = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
    res = surf_renderer_b200.render_splats_along_ray(sc, samples=2, normal_estimation_method='plane')
    assert lib().surf_last_launch_count() == 3          # k_splat_setup, k_splat_normals, k_splat_forward
    res['image'].sum().backward()
    torch.cuda.synchronize()
    assert lib().surf_last_launch_count() == 6          # setup, normals, backward, normals_backward, src_finalize, finalize
    assert res['image'].shape == (64, 64, 3)


@pytest.mark.skipif(not os.path.isdir('/root/reference/diffrend'), reason='reference tree not present')
def test_z_to_pcl_cc_matches_live_reference():
    """renderer.py:484-534 (imported by the GAN trainer next to the renderers): pure tensor programs, bit-exact."""
    import sys
    sys.path.insert(0, '/root/reference')
    from diffrend.torch.renderer import z_to_pcl_CC as ref_fn, z_to_pcl_CC_batched as ref_batched
    import surf_renderer_b200
    cam = {'viewport': [0, 0, 37, 23], 'fovy': float(np.deg2rad(33.)), 'focal_length': 0.7}
    g = torch.Generator().manual_seed(2)
    z = -(torch.rand(37 * 23, generator=g) * 4 + 0.5)
    z[::7] = 0.3                                             # behind the camera: clamped to 0
    za = z.clone().requires_grad_(True)
    zb = z.clone().requires_grad_(True)
    a, b = ref_fn(za, cam), surf_renderer_b200.z_to_pcl_CC(zb, cam)
    assert torch.equal(a, b)
    (a * torch.arange(3.)).sum().backward()
    (b * torch.arange(3.)).sum().backward()
    assert torch.equal(za.grad, zb.grad)
    zz = torch.stack((z, z * 0.5, z * 2))
    assert torch.equal(ref_batched(zz, cam), surf_renderer_b200.z_to_pcl_CC_batched(zz, cam))
