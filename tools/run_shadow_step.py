"""N forward renders of config E with shadow rays (100K splats, 1024x1024, 3 lights) - the command profiled under ncu
for k_intersect_shadow (profiles/r1_shadow_config_e_ncu_full.json)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import scene_io, surf_renderer_b200
from surf_renderer_b200 import scenes as synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sc = scene_io.clone_scene(synth.config_e(), device='cuda')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for i in range(n):
        if i == n - 1:
            e0.record()
        surf_renderer_b200.render(sc, shadow=True)
e1.record()
torch.cuda.synchronize()
print('last forward with shadows %.3f ms' % e0.elapsed_time(e1))
