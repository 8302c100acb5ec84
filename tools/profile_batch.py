"""Where the time of a config-D batch (64 scenes x 5000 splats, 128x128, fwd+bwd; BASELINE configs[3]) goes, for the
three ways of rendering it: a Python loop of render(), render_batch(list of scene dicts) and render_batch(one dict of
stacked tensors).  Host = wall clock until the call returns; host+gpu = until the stream drains; GPU = CUDA events.
Diagnostic tool (tools/trace_batch.py prints the per-kernel timeline)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import scene_io, surf_renderer_b200
from surf_renderer_b200 import scenes as synth
from surf_renderer_b200.renderer import _stack_scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
GRAD_FIELDS = ('pos', 'normal')          # what a generator would learn: splat positions and normals


def leafy(sc):
    for f in GRAD_FIELDS:
        sc['objects']['disk'][f].requires_grad_(True)
    return sc


batch = [leafy(scene_io.clone_scene(synth.config_d_scene(i), device='cuda')) for i in range(n)]
stacked = leafy(_stack_scenes([scene_io.clone_scene(synth.config_d_scene(i), device='cuda') for i in range(n)]))


def section(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    host, total, gpu = [], [], []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0.record(); fn(); e1.record()
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        host.append(t1 - t0); total.append(t2 - t0); gpu.append(e0.elapsed_time(e1))
    print('%-44s host %7.3f ms   host+gpu %7.3f ms   GPU %7.3f ms' % (name, min(host) * 1e3, min(total) * 1e3, min(gpu)), flush=True)
    return min(total) * 1e3


def fb_loop():
    rs = [surf_renderer_b200.render(sc, double_sided=True) for sc in batch]
    sum(r['image'].sum() for r in rs).backward()


def fb_list():
    rs = surf_renderer_b200.render_batch(batch, double_sided=True)
    sum(r['image'].sum() for r in rs).backward()


def fb_stacked():
    surf_renderer_b200.render_batch(stacked, double_sided=True)['image'].sum().backward()


with torch.no_grad():
    section('forward: loop of render()', lambda: [surf_renderer_b200.render(sc, double_sided=True) for sc in batch])
    section('forward: render_batch(list)', lambda: surf_renderer_b200.render_batch(batch, double_sided=True))
    section('forward: render_batch(stacked dict)', lambda: surf_renderer_b200.render_batch(stacked, double_sided=True))
section('fwd+bwd: loop of render()', fb_loop)
section('fwd+bwd: render_batch(list)', fb_list)
t = section('fwd+bwd: render_batch(stacked dict)', fb_stacked)
print('stacked: %.3e ray-splat tests/s' % (n * 5000 * 128 * 128 / (t * 1e-3)))
