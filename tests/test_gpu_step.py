"""GPU tests (-m gpu) of the fused inverse-rendering step (surf_step_mse / MSEStep; reference loop body:
diffrend/torch/test_optimization.py:100-125), of scenes beyond the old fixed capacities (many lights / materials), and
of the device-side index guard."""
import numpy as np
import pytest
import torch

import parity
import scene_io
from oracle import torch_oracle

pytestmark = pytest.mark.gpu


def _leaves(sc):
    return scene_io.grad_leaves(sc)


def _mse_reference(scene, target, params):
    """loss and gradients of mean((image - target)^2) through the oracle's autograd"""
    osc = scene_io.clone_scene(scene, requires_grad=True)
    res = torch_oracle.render(osc, **params)
    loss = ((res['image'] - target) ** 2).mean()
    loss.backward()
    return float(loss), {k: v.grad for k, v in _leaves(osc).items() if v.grad is not None}, res


@pytest.mark.parametrize('which', ['mixed', 'splats', 'torus_like'])
def test_mse_step_matches_oracle_autograd(which):
    import surf_renderer_b200
    from surf_renderer_b200 import scenes as synth
    if which == 'mixed':
        scene = synth.random_mixed_scene(17, width=72, height=56, n_disk=30, n_tri=20, n_sphere=0)
        params = {'double_sided': True}
    elif which == 'splats':
        scene = synth.config_e(m=3000, width=96, height=96, radius=0.03)
        params = {}
    else:
        scene = synth.random_mixed_scene(5, width=64, height=64, n_disk=0, n_plane=0, n_sphere=0, n_tri=60)
        params = {'double_sided': True, 'use_quartic': True}
    g = torch.Generator().manual_seed(3)
    H, W = scene['camera']['viewport'][3], scene['camera']['viewport'][2]
    target = torch.rand(H, W, 3, generator=g)
    # drop eps-tie / kink pixels from the loss on both sides by making the target equal the render there
    ref0 = torch_oracle.render(scene_io.clone_scene(scene), **params)
    ref_np = {k: v.detach() for k, v in ref0.items() if isinstance(v, torch.Tensor)}
    cand0 = surf_renderer_b200.render(scene_io.clone_scene(scene, device='cuda'), **params)
    rep = parity.compare_forward({k: (v.detach().cpu() if isinstance(v, torch.Tensor) else v) for k, v in cand0.items()}, ref_np, scene)
    kink = parity.kink_mask(scene, ref_np, params) & (ref0['depth'].detach().reshape(-1).numpy() <= scene['camera']['far'])
    bad = torch.tensor(~rep['good_mask'] | kink).view(H, W)
    target_o = torch.where(bad[..., None], ref0['image'].detach(), target)
    target_c = torch.where(bad[..., None], cand0['image'].detach().cpu(), target)
    ref_loss, ref_grads, _ = _mse_reference(scene, target_o, params)

    sc = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
    plan = surf_renderer_b200.MSEStep(sc, target_c.cuda(), **params)
    loss = plan()
    torch.cuda.synchronize()
    assert abs(float(loss) - ref_loss) <= 1e-5 * abs(ref_loss) + 1e-9
    lv = _leaves(sc)
    parity.compare_grads({k: lv[k].grad.cpu() for k in ref_grads}, ref_grads)
    # a second call reproduces the first (the packed buffer is re-zeroed, nothing accumulates across steps)
    g0 = {k: lv[k].grad.clone() for k in ref_grads}
    loss2 = plan()
    torch.cuda.synchronize()
    assert abs(float(loss2) - float(loss)) <= 1e-6 * abs(float(loss))
    for k in ref_grads:
        assert torch.allclose(lv[k].grad, g0[k], rtol=1e-4, atol=1e-6 * float(g0[k].abs().max()))
    assert plan.launches >= 6          # setup, raygen, prep, intersect, shade, backward, finalize (memsets not counted)


def test_mse_step_equals_the_autograd_path_of_this_library():
    """render() -> MSE -> backward() through the autograd.Function and MSEStep run the same kernels"""
    import surf_renderer_b200
    from surf_renderer_b200 import scenes as synth
    scene = synth.config_e(m=20000, width=256, height=256, radius=0.012)
    target = surf_renderer_b200.render(scene_io.clone_scene(synth.config_e_target_scene(scene), device='cuda'))['image'].detach()
    a = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
    res = surf_renderer_b200.render(a)
    loss_a = ((res['image'] - target) ** 2).mean()
    loss_a.backward()
    b = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
    plan = surf_renderer_b200.MSEStep(b, target)
    loss_b = plan()
    torch.cuda.synchronize()
    assert torch.equal(plan.image.view(256, 256, 3), res['image'])
    assert abs(float(loss_a) - float(loss_b)) <= 2e-6 * abs(float(loss_a))
    la, lb = _leaves(a), _leaves(b)
    for k in la:
        if la[k].grad is None:
            continue
        scale = float(la[k].grad.abs().max())
        assert torch.allclose(lb[k].grad, la[k].grad, rtol=1e-4, atol=2e-6 * scale), k


@pytest.mark.parametrize('variant', ['shadow', 'ortho', 'mode3', 'spheres'])
def test_mse_step_option_space_equals_the_autograd_path(variant):
    """the fused step honours the same options as render(): shadow rays (visibility from the forward workspace),
    orthographic frames, the screen-space intersection kernel, sphere primitives"""
    import surf_renderer_b200
    from surf_renderer_b200 import scenes as synth
    params = {'double_sided': True}
    if variant == 'shadow':
        scene = synth.random_mixed_scene(31, width=64, height=48, n_disk=40, n_tri=20, n_sphere=0)
        params['shadow'] = True
    elif variant == 'ortho':
        scene = synth.random_mixed_scene(32, width=56, height=40, n_disk=30, n_tri=20, n_sphere=0, proj='orthographic')
    elif variant == 'mode3':
        scene = synth.random_mixed_scene(33, width=130, height=17, n_disk=40, n_tri=20, n_sphere=0)
        params['_math_mode'] = 3
    else:
        scene = synth.random_mixed_scene(34, width=64, height=48, n_disk=20, n_tri=10, n_sphere=4)
    H, W = scene['camera']['viewport'][3], scene['camera']['viewport'][2]
    target = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(2)).cuda()
    a = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
    res = surf_renderer_b200.render(a, **params)
    loss_a = ((res['image'] - target) ** 2).mean()
    loss_a.backward()
    b = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
    plan = surf_renderer_b200.MSEStep(b, target, **params)
    loss_b = plan()
    torch.cuda.synchronize()
    assert torch.equal(plan.image.view(H, W, 3), res['image'])
    assert abs(float(loss_a) - float(loss_b)) <= 2e-6 * abs(float(loss_a))
    la, lb = _leaves(a), _leaves(b)
    checked = 0
    for k in la:
        if la[k].grad is None:
            continue
        scale = float(la[k].grad.abs().max())
        assert torch.allclose(lb[k].grad, la[k].grad, rtol=1e-4, atol=3e-6 * max(scale, 1e-30)), k
        checked += 1
    assert checked >= 5


def test_mse_step_with_adam_replays_from_a_cuda_graph():
    import surf_renderer_b200
    from surf_renderer_b200 import scenes as synth
    scene = synth.config_e(m=4000, width=128, height=128, radius=0.03)
    target = surf_renderer_b200.render(scene_io.clone_scene(synth.config_e_target_scene(scene, jitter=0.01), device='cuda'))['image'].detach()

    def make():
        sc = scene_io.clone_scene(scene, device='cuda')
        pos = sc['objects']['disk']['pos'].requires_grad_(True)
        plan = surf_renderer_b200.MSEStep(sc, target)
        opt = torch.optim.Adam([pos], lr=1e-3, capturable=True)

        def step():
            loss = plan()
            opt.step()
            return loss
        return pos, step

    pos_e, step_e = make()
    losses_e = [float(step_e()) for _ in range(8)]
    pos_g, step_g = make()
    graphed = surf_renderer_b200.GraphedStep(step_g, warmup=3)
    losses_g = [float(graphed()) for _ in range(5)]
    torch.cuda.synchronize()
    assert losses_e[-1] < losses_e[0]
    assert np.allclose(losses_g, losses_e[3:8], rtol=2e-4)
    assert torch.allclose(pos_g, pos_e, rtol=1e-4, atol=1e-6)


def test_packed_adam_matches_torch_adam():
    """surf_adam_step (one kernel over MSEStep's packed gradients) follows torch.optim.Adam step for step"""
    import surf_renderer_b200
    from surf_renderer_b200 import scenes as synth
    scene = synth.config_e(m=3000, width=96, height=96, radius=0.03)
    target = surf_renderer_b200.render(scene_io.clone_scene(synth.config_e_target_scene(scene, jitter=0.01), device='cuda'))['image'].detach()

    def run(packed):
        sc = scene_io.clone_scene(scene, device='cuda')
        leaves = [sc['objects']['disk']['pos'], sc['objects']['disk']['normal'], sc['materials']['albedo'], sc['lights']['pos']]
        for t in leaves:
            t.requires_grad_(True)
        plan = surf_renderer_b200.MSEStep(sc, target)
        opt = surf_renderer_b200.PackedAdam(plan, lr=2e-3) if packed else torch.optim.Adam(leaves, lr=2e-3)
        losses = []
        for _ in range(6):
            losses.append(float(plan()))
            opt.step()
        torch.cuda.synchronize()
        return leaves, losses

    la, loss_a = run(True)
    lb, loss_b = run(False)
    assert np.allclose(loss_a, loss_b, rtol=1e-4)
    for a, b in zip(la, lb):
        assert torch.allclose(a, b, rtol=1e-5, atol=2e-6), float((a - b).abs().max())


def test_many_lights_and_materials_beyond_the_old_capacities():
    """24 lights and 200 materials: 2 x 600 + 2 x 72 + ... = more than the 512 shared accumulator slots of the backward
    (the overflow goes straight to the leaves), more than 16 lights with shadow rays."""
    import surf_renderer_b200
    from surf_renderer_b200 import scenes as synth
    g = torch.Generator().manual_seed(11)
    scene = synth.random_mixed_scene(23, width=64, height=48, n_disk=120, n_plane=1, n_sphere=0, n_tri=80)
    K, L = 200, 24
    scene['materials'] = {'albedo': torch.rand(K, 3, generator=g) * 0.8 + 0.1,
                          'coeffs': torch.cat((torch.rand(K, 2, generator=g), torch.rand(K, 1, generator=g) * 20 + 1), dim=1)}
    for kind, prim in scene['objects'].items():
        cnt = prim['material_idx'].shape[0]
        prim['material_idx'] = torch.randint(0, K, (cnt,), generator=g)
    ang = torch.rand(L, generator=g) * 6.28
    scene['lights'] = {'pos': torch.stack((8 * torch.cos(ang), 4 + 4 * torch.rand(L, generator=g), 8 * torch.sin(ang), torch.ones(L)), dim=1),
                       'color_idx': torch.randint(0, scene['colors'].shape[0], (L,), generator=g),
                       'attenuation': torch.tensor([[1.0, 0.02, 0.002]]).repeat(L, 1) * (0.5 + torch.rand(L, 1, generator=g)),
                       'ambient': torch.tensor([0.002, 0.002, 0.002])}
    for params in ({'double_sided': True}, {'double_sided': True, 'shadow': True}):
        sc = scene_io.clone_scene(scene, device='cuda', requires_grad=True)
        res = surf_renderer_b200.render(sc, **params)
        osc = scene_io.clone_scene(scene, requires_grad=True)
        ref = torch_oracle.render(osc, **params)
        ref_np = {k: v.detach() for k, v in ref.items() if isinstance(v, torch.Tensor)}
        rep = parity.compare_forward({k: (v.detach().cpu() if isinstance(v, torch.Tensor) else v) for k, v in res.items()}, ref_np, scene)
        H, W = ref['depth'].shape
        far = scene['camera']['far']
        kink = parity.kink_mask(scene, ref_np, params) & (ref['depth'].detach().reshape(-1).numpy() <= far)
        good = torch.tensor(rep['good_mask'] & ~kink).view(H, W)
        w = scene_io.loss_weights((H, W), 5)
        for k in w:
            w[k] = w[k] * (good[..., None] if w[k].dim() == 3 else good)
        scene_io.weighted_loss(ref, w, far).backward()
        scene_io.weighted_loss(res, w, far).backward()
        lo, lg = _leaves(osc), _leaves(sc)
        names = [k for k in lo if lo[k].grad is not None]
        assert 'materials/albedo' in names and 'lights/pos' in names
        parity.compare_grads({k: lg[k].grad.cpu() for k in names}, {k: lo[k].grad for k in names})


def test_more_than_eight_primitive_sets_are_rejected_like_before_or_rendered():
    """the reference has four primitive kinds, so a scene dict can hold at most four sets; the ABI allows eight"""
    import surf_renderer_b200
    from surf_renderer_b200 import scenes as synth
    scene = synth.random_mixed_scene(3, width=40, height=32)
    assert len(scene['objects']) <= 4
    res = surf_renderer_b200.render(scene_io.clone_scene(scene, device='cuda'))
    assert res['image'].shape == (32, 40, 3)


def test_device_side_index_guard():
    """material_idx / color_idx held in DEVICE memory: clamped by the kernels, reported by _check_indices=True as the
    IndexError the reference's index_select raises; host-side index arrays raise up front"""
    import surf_renderer_b200
    from surf_renderer_b200 import scenes as synth
    scene = synth.config_e(m=500, width=48, height=48, radius=0.05)
    sc = scene_io.clone_scene(scene, device='cuda')
    good = surf_renderer_b200.render(sc, _check_indices=True)['image']
    bad = scene_io.clone_scene(scene, device='cuda')
    bad['objects']['disk']['material_idx'] = torch.full((500,), 7, dtype=torch.int64, device='cuda')
    out = surf_renderer_b200.render(bad)['image']           # clamped to the last material row, no fault
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    with pytest.raises(IndexError):
        surf_renderer_b200.render(bad, _check_indices=True)
    host_bad = scene_io.clone_scene(scene, device='cuda')
    host_bad['objects']['disk']['material_idx'] = [3] * 500
    with pytest.raises(IndexError):
        surf_renderer_b200.render(host_bad)
    host_bad2 = scene_io.clone_scene(scene, device='cuda')
    host_bad2['lights']['color_idx'] = np.full(3, -1)
    with pytest.raises(IndexError):
        surf_renderer_b200.render(host_bad2)
    assert torch.equal(good, surf_renderer_b200.render(sc)['image'])
