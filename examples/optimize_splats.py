"""Inverse rendering of a splat cloud, the loop of the reference's optimize_scene (diffrend/torch/test_optimization.py:
69-125): render -> MSE against a target image -> backward -> Adam.  The splats' normals and the material albedo start
perturbed and are recovered from the shading (splat positions also receive gradients - along their normals, through
depth and lighting - but hard-edged disks give no lateral signal, as in the reference).  The same step is run eagerly
and replayed from a CUDA graph (surf_renderer_b200.GraphedStep); both walk the same trajectory.

    python examples/optimize_splats.py [n_splats] [size] [iterations]
"""
import sys
import time

import torch

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import surf_renderer_b200 as surf                                   # noqa: E402
from surf_renderer_b200 import scenes as synth                      # noqa: E402
from surf_renderer_b200.scenes import clone_scene                   # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
size = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200

target_scene = clone_scene(synth.config_e(m=m, width=size, height=size, radius=0.03), device='cuda')
with torch.no_grad():
    target = surf.render(target_scene)['image'].clone()


def make_problem():
    scene = clone_scene(target_scene, device='cuda')
    disk = scene['objects']['disk']
    g = torch.Generator(device='cuda').manual_seed(0)
    disk['normal'] = (disk['normal'] + 0.3 * torch.randn(disk['normal'].shape, device='cuda', generator=g)).requires_grad_(True)
    scene['materials']['albedo'] = torch.tensor([[0.3, 0.7, 0.5]], device='cuda', requires_grad=True)
    params = [disk['normal'], scene['materials']['albedo']]
    for p in params:
        p.grad = torch.zeros_like(p)                                # static gradient buffers for the graph
    opt = torch.optim.Adam(params, lr=2e-2, capturable=True)

    def step():
        opt.zero_grad(set_to_none=False)
        loss = ((surf.render(scene)['image'] - target) ** 2).mean()
        loss.backward()
        opt.step()
        return loss
    return step


def run(fn, n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        loss = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, float(loss.detach())


eager = make_problem()
first = float(eager().detach())
ms_eager, loss_eager = run(eager, iters - 1)
graphed = surf.GraphedStep(make_problem(), warmup=1)                 # 1 eager step + capture, then replays
ms_graph, loss_graph = run(graphed, iters - 1)
print('%d splats, %dx%d, %d Adam iterations: loss %.6f -> %.6f' % (m, size, size, iters, first, loss_eager))
print('eager  %.3f ms/iteration' % ms_eager)
print('graph  %.3f ms/iteration   (final loss %.6f)' % (ms_graph, loss_graph))
