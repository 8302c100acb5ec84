"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard a frame into row bands, all-gather the
outputs and all-reduce packed gradients (surf_renderer_b200/dist.py).  The CUDA renderer cannot run here, so the
band renderer is the CPU oracle (pixel_subset) - the thing under test is the sharding / collective plumbing."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_render_flat(scene, pixel_range, **params):
    from oracle import torch_oracle
    p0, p1 = pixel_range
    vp = scene['camera']['viewport']
    W, H = vp[2] - vp[0], vp[3] - vp[1]
    res = torch_oracle.render(scene, pixel_subset=torch.arange(p0, p1), **params)
    n = p1 - p0
    return (res['image'].reshape(n, 3), res['depth'].reshape(n), res['normal'].reshape(n, 3), res['pos'].reshape(n, 3),
            res['nearest'].reshape(n), res['ray_dir']), (H, W)


def _oracle_render_batch(scene, **params):
    """stand-in for render_batch(stacked dict): the oracle, scene by scene"""
    from oracle import torch_oracle
    from surf_renderer_b200.marshal import batched_scene_size, select_scenes
    rs = [torch_oracle.render(select_scenes(scene, b), **params) for b in range(batched_scene_size(scene))]
    return {k: torch.stack([r[k] for r in rs]) for k in ('image', 'depth', 'normal', 'pos', 'nearest')}


def _check_scene_sharding(rank, world):
    """scenes over ranks (config D): 5 stacked scenes on 2 ranks (blocks of 3 and 2), shared lights"""
    from surf_renderer_b200 import dist as sdist, scenes as synth
    from surf_renderer_b200.renderer import _stack_scenes

    def build():
        lights = torch.tensor([[2., 3., 4., 1.], [-3., 1., 5., 1.]], requires_grad=True)
        parts = []
        for i in range(5):
            sc = synth.config_d_scene(i, m=60, width=12, height=10, radius=0.12)
            sc['lights'] = {'pos': lights, 'color_idx': sc['lights']['color_idx'][:2],
                            'attenuation': sc['lights']['attenuation'][:2], 'ambient': sc['lights']['ambient']}
            parts.append(sc)
        st = _stack_scenes(parts)
        st['objects']['disk']['pos'] = st['objects']['disk']['pos'].detach().requires_grad_(True)
        return st, lights
    w = torch.rand(5, 10, 12, 3, generator=torch.Generator().manual_seed(3))
    st, lights = build()
    res = sdist.render_batch_sharded(st, render_batch_fn=_oracle_render_batch, double_sided=True)
    (res['image'] * w).sum().backward()
    sdist.allreduce_gradients([st['objects']['disk']['pos'], lights])
    st1, lights1 = build()
    ref = _oracle_render_batch(st1, double_sided=True)
    (ref['image'] * w).sum().backward()
    b0, b1 = res['block']
    ok = res['image'].shape == (5, 10, 12, 3) and torch.equal(res['image'], ref['image'])
    ok = ok and torch.equal(res['nearest'], ref['nearest']) and res['pos'].shape[0] == b1 - b0
    ok = ok and torch.allclose(st['objects']['disk']['pos'].grad, st1['objects']['disk']['pos'].grad, rtol=1e-4, atol=1e-7)
    ok = ok and torch.allclose(lights.grad, lights1.grad, rtol=1e-4, atol=1e-6)
    return bool(ok)


def _check_grad_bucket(rank, world):
    """GradBucket: the leaves' .grad are views of one flat buffer; autograd accumulates into them in place and ONE
    in-place all-reduce sums them over ranks - same numbers as allreduce_gradients' pack / reduce / copy-back"""
    from surf_renderer_b200 import dist as sdist
    g = torch.Generator().manual_seed(7)
    a0, b0 = torch.rand(5, 3, generator=g), torch.rand(4, generator=g)
    a, b = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    bucket = sdist.GradBucket([a, b])
    ((a * (rank + 1)).sum() + (b ** 2).sum() * (rank + 2)).backward()
    ok = a.grad.data_ptr() == bucket.views[0].data_ptr() and b.grad.data_ptr() == bucket.views[1].data_ptr()
    bucket.all_reduce()
    exp_a = torch.full((5, 3), float(sum(r + 1 for r in range(world))))
    exp_b = 2 * b0 * float(sum(r + 2 for r in range(world)))
    ok = ok and torch.allclose(a.grad, exp_a) and torch.allclose(b.grad, exp_b)
    bucket.zero_()
    ok = ok and float(a.grad.abs().sum()) == 0.0 and a.grad.data_ptr() == bucket.views[0].data_ptr()
    return bool(ok)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import scene_io
    from oracle import torch_oracle
    from surf_renderer_b200 import dist as sdist, scenes as synth
    scene = synth.random_mixed_scene(11, width=31, height=17, n_sphere=0)      # 527 pixels: uneven split
    sc = scene_io.clone_scene(scene, requires_grad=True)
    res = sdist.render_bands(sc, render_flat_fn=_oracle_render_flat, gather=('image', 'depth', 'nearest'))
    g = torch.Generator().manual_seed(0)
    w = torch.rand(17, 31, 3, generator=g)
    loss = (res['image'] * w).sum() + (res['depth'].clamp(max=50) * 0.1).sum()
    loss.backward()
    leaves = list(scene_io.grad_leaves(sc).values())
    nbytes = sdist.allreduce_gradients([t for t in leaves if t.requires_grad])
    # single-process reference
    sc1 = scene_io.clone_scene(scene, requires_grad=True)
    ref = torch_oracle.render(sc1)
    loss1 = (ref['image'] * w).sum() + (ref['depth'].clamp(max=50) * 0.1).sum()
    loss1.backward()
    ok = torch.equal(res['image'], ref['image']) and torch.equal(res['nearest'], ref['nearest']) \
        and torch.equal(res['depth'], ref['depth'])
    l0, l1 = scene_io.grad_leaves(sc), scene_io.grad_leaves(sc1)
    for k in l1:
        if l1[k].grad is None:
            continue
        ok = ok and l0[k].grad is not None and torch.allclose(l0[k].grad, l1[k].grad, rtol=1e-4, atol=1e-6)
    # scenes-over-ranks helpers
    mine = sdist.shard_scenes(5, rank, world)
    stack = torch.tensor([[float(i)] for i in mine])
    full = sdist.gather_scene_outputs(stack, 5)
    ok = ok and torch.equal(full.reshape(-1), torch.arange(5.0)) and nbytes > 0
    ok = ok and _check_scene_sharding(rank, world)
    ok = ok and _check_grad_bucket(rank, world)
    open(os.path.join(tmp, 'ok%d' % rank), 'w').write('1' if ok else '0')
    dist.destroy_process_group()


def test_band_sharding_two_ranks_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / ('ok%d' % r)).read() == '1'


def test_band_range_partitions_exactly():
    from surf_renderer_b200.dist import band_range
    for n in (1, 7, 1024 * 1024, 527):
        for world in (1, 2, 3, 8):
            spans = [band_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
