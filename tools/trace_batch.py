"""Kernel timeline of one stacked config-D batch (torch profiler chrome trace): per kernel name, the summed
duration, the wall-clock union and the resulting average concurrency across the library's internal streams."""
import json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import scene_io, surf_renderer_b200
from surf_renderer_b200 import scenes as synth
from surf_renderer_b200.renderer import _stack_scenes
from torch.profiler import profile, ProfilerActivity

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
stacked = _stack_scenes([scene_io.clone_scene(synth.config_d_scene(i), device='cuda') for i in range(n)])
for f in ('pos', 'normal'):
    stacked['objects']['disk'][f].requires_grad_(True)


def step():
    r = surf_renderer_b200.render_batch(stacked, double_sided=True)
    r['image'].sum().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
path = os.path.join(ROOT, 'gpurun_out', 'batch_trace.json')
os.makedirs(os.path.dirname(path), exist_ok=True)
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')]
t0 = min(e['ts'] for e in ev)
by = collections.defaultdict(list)
for e in ev:
    by[e['name'].split('(')[0][:40]].append((e['ts'] - t0, e['ts'] - t0 + e['dur'], e['args'].get('stream')))
print('%-42s %5s %9s %9s %6s %s' % ('kernel', 'n', 'sum us', 'union us', 'conc', 'streams'))
for name, iv in sorted(by.items(), key=lambda kv: -sum(b - a for a, b, _ in kv[1])):
    iv.sort()
    union, cur_a, cur_b = 0.0, None, None
    for a, b, _ in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                union += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    union += cur_b - cur_a
    tot = sum(b - a for a, b, _ in iv)
    print('%-42s %5d %9.1f %9.1f %6.2f %d   first %.0f last %.0f' % (name, len(iv), tot, union, tot / union, len({s for _, _, s in iv}), iv[0][0], iv[-1][1]))
os.remove(path)
