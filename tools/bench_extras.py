"""Timings of the 'next' rows: shadow rays, orthographic camera, render_splats_along_ray.  CUDA events."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import scene_io, surf_renderer_b200
from surf_renderer_b200 import scenes as synth


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
for name, scene in (('splats5k_256_7lights', synth.splat_scene(*synth.synthetic_sphere_splats(5000, 3), 0.02, 256, 256, 14., 1., (0., 0., 5., 1.))),
                    ('config_e_100k_1024_3lights', synth.config_e())):
    sc = scene_io.clone_scene(scene, device='cuda')
    with torch.no_grad():
        t0 = timed(lambda: surf_renderer_b200.render(sc))
        t1 = timed(lambda: surf_renderer_b200.render(sc, shadow=True), reps=3, warm=1)
    M = sc['objects']['disk']['pos'].shape[0]; N = sc['camera']['viewport'][2] * sc['camera']['viewport'][3]; L = sc['lights']['pos'].shape[0]
    out[name] = {'forward_ms': t0, 'forward_shadow_ms': t1, 'shadow_tests_per_s': M * N * L / ((t1 - t0) * 1e-3)}
    print(name, out[name], flush=True)

ortho = synth.random_mixed_scene(34, width=512, height=512, n_disk=20000, n_tri=0, n_sphere=0, n_plane=0, proj='orthographic')
sc = scene_io.clone_scene(ortho, device='cuda')
with torch.no_grad():
    t = timed(lambda: surf_renderer_b200.render(sc))
out['ortho_20k_512'] = {'forward_ms': t, 'tests_per_s': 20000 * 512 * 512 / (t * 1e-3)}
print('ortho', out['ortho_20k_512'], flush=True)

from make_golden_along_ray_scene import along_ray_scene
for size in (128, 1024):
    sc = scene_io.clone_scene(along_ray_scene(5, size, size, mats=1), device='cuda', requires_grad=True)

    def fb():
        r = surf_renderer_b200.render_splats_along_ray(sc)
        (r['image'].sum() + r['depth'].sum()).backward()
    with torch.no_grad():
        tf = timed(lambda: surf_renderer_b200.render_splats_along_ray(sc))
    tfb = timed(fb)
    n = size * size
    out['along_ray_%d' % size] = {'forward_ms': tf, 'fwd_bwd_ms': tfb, 'forward_gbs': n * 60 / (tf * 1e-3) / 1e9}
    print('along_ray', size, out['along_ray_%d' % size], flush=True)
# config D: 64 scenes x 5000 splats at 128x128, double sided, fwd + bwd (BASELINE configs[3]); loop of render() vs render_batch()
batch = [scene_io.clone_scene(synth.config_d_scene(i), device='cuda', requires_grad=True) for i in range(64)]


def loop_fb():
    rs = [surf_renderer_b200.render(sc, double_sided=True) for sc in batch]
    sum(r['image'].sum() for r in rs).backward()


def batch_fb():
    rs = surf_renderer_b200.render_batch(batch, double_sided=True)
    sum(r['image'].sum() for r in rs).backward()


from surf_renderer_b200.renderer import _stack_scenes
stacked = _stack_scenes([scene_io.clone_scene(synth.config_d_scene(i), device='cuda') for i in range(64)])
for f in ('pos', 'normal'):
    stacked['objects']['disk'][f].requires_grad_(True)


def stacked_fb():
    surf_renderer_b200.render_batch(stacked, double_sided=True)['image'].sum().backward()


t_loop = timed(loop_fb, reps=3, warm=1)
t_batch = timed(batch_fb, reps=3, warm=1)
t_stacked = timed(stacked_fb, reps=5, warm=2)
tests = 64 * 5000 * 128 * 128
out['config_d_64x5000_128'] = {'loop_fwd_bwd_ms': t_loop, 'batch_list_fwd_bwd_ms': t_batch, 'batch_stacked_fwd_bwd_ms': t_stacked,
                               'stacked_tests_per_s': tests / (t_stacked * 1e-3),
                               'note': 'loop/list: every float leaf of every scene requires grad; stacked: pos + normal'}
print('config_d', out['config_d_64x5000_128'], flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'bench_extras.json'), 'w'), indent=1)
