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
