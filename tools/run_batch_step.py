"""N fwd+bwd steps of the stacked config-D batch (64 scenes x 5000 splats, 128x128) - the command profiled under ncu
for the *_batch kernels (profiles/r1_batch_*)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import scene_io, surf_renderer_b200
from surf_renderer_b200 import scenes as synth
from surf_renderer_b200.renderer import _stack_scenes
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
st = _stack_scenes([scene_io.clone_scene(synth.config_d_scene(i), device='cuda') for i in range(64)])
for f in ('pos', 'normal'):
    st['objects']['disk'][f].requires_grad_(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(steps):
    if i == steps - 1:
        e0.record()
    surf_renderer_b200.render_batch(st, double_sided=True)['image'].sum().backward()
e1.record()
torch.cuda.synchronize()
print('last step %.3f ms' % e0.elapsed_time(e1))
