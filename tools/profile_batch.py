"""Where the time of a config-D batch (64 scenes x 5000 splats, 128x128, fwd+bwd) goes: host marshalling vs library
call vs GPU execution.  Wall clock with synchronisation around each section (diagnostic, not a bench number)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import scene_io, surf_renderer_b200
from surf_renderer_b200 import scenes as synth
from surf_renderer_b200.marshal import Marshalled
from surf_renderer_b200._lib import lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
batch = [scene_io.clone_scene(synth.config_d_scene(i), device='cuda', requires_grad=True) for i in range(n)]
dev = torch.device('cuda', 0)


def section(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn()
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        t.append((t1 - t0, t2 - t0))
    host = min(x[0] for x in t) * 1e3
    total = min(x[1] for x in t) * 1e3
    print('%-34s host %7.3f ms   host+gpu %7.3f ms' % (name, host, total), flush=True)
    return r


section('marshal x%d' % n, lambda: [Marshalled(sc, dev) for sc in batch])
with torch.no_grad():
    section('render_batch forward (no_grad)', lambda: surf_renderer_b200.render_batch(batch, double_sided=True))
rs = section('render_batch forward (grad)', lambda: surf_renderer_b200.render_batch(batch, double_sided=True))
loss = section('loss = sum of image sums', lambda: sum(r['image'].sum() for r in rs))


def fb():
    rs = surf_renderer_b200.render_batch(batch, double_sided=True)
    sum(r['image'].sum() for r in rs).backward()


section('fwd + loss + bwd', fb)
lib().surf_set_kernel_timing(1)
fb(); torch.cuda.synchronize()
print('launches in last call:', lib().surf_last_launch_count())
for k in range(6):
    print('  kernel slot %d: %.4f ms' % (k, lib().surf_last_kernel_ms(k)))
