"""render_splats_along_ray timings: one frame, the per-element loop of gan.py:563-597 over a batch of
64 frames at 128x128, and the same batch through render_splats_along_ray_batch (one call per direction) - fwd+bwd,
with the GAN's call (samples=1|2, normal_estimation_method='plane') and with given normals."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch                      # noqa: E402
import scene_io                   # noqa: E402
import surf_renderer_b200         # noqa: E402
from make_golden_along_ray_scene import along_ray_scene   # noqa: E402

B, S = 64, 128


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows = []
base = scene_io.clone_scene(along_ray_scene(7, S, S, mats=1, smooth_z=True), device='cuda')
zs = torch.stack([scene_io.clone_scene(along_ray_scene(100 + b, S, S, mats=1, smooth_z=True))['objects']['disk']['pos'] for b in range(B)]).cuda()
zs = (zs + 0.05 * torch.randn_like(zs)).requires_grad_(True)
ns = torch.nn.functional.normalize(torch.randn(B, S * S, 3, device='cuda') + torch.tensor([0., 0., 2.], device='cuda'), dim=-1).requires_grad_(True)
eyes = torch.stack([torch.tensor([0.3 * (b % 5), 0.1 * (b % 7), 5.0, 1.0]) for b in range(B)]).cuda()
for name, params, given in (('given normals', {}, True), ("estimated normals ('plane')", {'normal_estimation_method': 'plane'}, False),
                            ("estimated normals ('plane'), samples=2", {'normal_estimation_method': 'plane', 'samples': 2}, False)):
    def scene_of(z, n, eye):
        sc = dict(base)
        sc['objects'] = {'disk': dict(base['objects']['disk'])}
        sc['objects']['disk']['pos'] = z
        if given:
            sc['objects']['disk']['normal'] = n
        else:
            sc['objects']['disk'].pop('normal', None)
        sc['camera'] = dict(base['camera'])
        sc['camera']['eye'] = eye
        return sc

    def one():
        zs.grad = None
        r = surf_renderer_b200.render_splats_along_ray(scene_of(zs[0], ns[0], eyes[0]), **params)
        r['image'].sum().backward()

    def loop():
        zs.grad = None
        loss = 0
        for b in range(B):
            r = surf_renderer_b200.render_splats_along_ray(scene_of(zs[b], ns[b], eyes[b]), **params)
            loss = loss + r['image'].sum()
        loss.backward()

    def batch():
        zs.grad = None
        r = surf_renderer_b200.render_splats_along_ray_batch(scene_of(zs, ns, eyes), **params)
        r['image'].sum().backward()

    row = {'case': name, 'frame': '%dx%d' % (S, S), 'batch': B, 'one_frame_fwd_bwd_ms': timed(one), 'loop_of_64_fwd_bwd_ms': timed(loop, 5),
           'batch_call_fwd_bwd_ms': timed(batch)}
    print(json.dumps(row), flush=True)
    rows.append(row)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'r2_bench_along_ray.json'), 'w'), indent=1)
