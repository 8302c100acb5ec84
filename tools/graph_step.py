"""An inverse-rendering step (render fwd + loss + bwd + Adam) captured ONCE into a CUDA graph and replayed: the
library launches everything asynchronously on the caller's stream with no hidden synchronisation, so torch.cuda.graph
can capture it as is.  Reports eager vs replay time per step on config B (bunny 256x256) and checks that replaying
gives the same parameters as stepping eagerly."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import scene_io, surf_renderer_b200


def make(size):
    scene, _, _, _, _ = scene_io.load_case(os.path.join(ROOT, 'tests', 'golden', 'b_bunny_48.npz'))
    scene['camera']['viewport'] = [0, 0, size, size]
    sc = scene_io.clone_scene(scene, device='cuda')
    pos = sc['objects']['disk']['pos']
    target = surf_renderer_b200.render(sc)['image'].detach().clone()
    sc['objects']['disk']['pos'] = (pos + 0.002 * torch.randn(pos.shape, device='cuda', generator=torch.Generator(device='cuda').manual_seed(1))).requires_grad_(True)
    opt = torch.optim.Adam([sc['objects']['disk']['pos']], lr=1e-4, capturable=True)
    return sc, target, opt


def step(sc, target, opt):
    opt.zero_grad(set_to_none=False)
    loss = ((surf_renderer_b200.render(sc)['image'] - target) ** 2).mean()
    loss.backward()
    opt.step()
    return loss


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
for size in (64, 256):
    sc, target, opt = make(size)
    sc['objects']['disk']['pos'].grad = torch.zeros_like(sc['objects']['disk']['pos'])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step(sc, target, opt)
    torch.cuda.current_stream().wait_stream(s)
    eager_ms = timed(lambda: step(sc, target, opt), 50)

    sc2, target2, opt2 = make(size)
    sc2['objects']['disk']['pos'].grad = torch.zeros_like(sc2['objects']['disk']['pos'])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step(sc2, target2, opt2)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        static_loss = step(sc2, target2, opt2)
    graph_ms = timed(g.replay, 50)
    # same trajectory: 3 warm-up + 50 timed (+1 capture run does not execute) steps on both
    sc3, target3, opt3 = make(size)
    sc3['objects']['disk']['pos'].grad = torch.zeros_like(sc3['objects']['disk']['pos'])
    for _ in range(53):
        step(sc3, target3, opt3)
    diff = float((sc3['objects']['disk']['pos'] - sc2['objects']['disk']['pos']).detach().abs().max())
    out['bunny_%d' % size] = {'eager_ms_per_step': eager_ms, 'graph_replay_ms_per_step': graph_ms, 'max_param_diff_vs_eager': diff,
                              'loss': float(static_loss)}
    print(size, out['bunny_%d' % size], flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'graph_step.json'), 'w'), indent=1)
