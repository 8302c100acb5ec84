"""How much of k_intersect is the exact narrow phase on dense scenes: config B (bunny 256x256) with the splat radii
scaled down (fewer filter candidates, same streaming work)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import scene_io, surf_renderer_b200
from surf_renderer_b200._lib import lib
scene, _, _, _, _ = scene_io.load_case(os.path.join(ROOT, 'tests', 'golden', 'b_bunny_48.npz'))
scene['camera']['viewport'] = [0, 0, 256, 256]
L = lib()
for scale in (1.0, 0.25, 0.001):
    sc = scene_io.clone_scene(scene, device='cuda')
    sc['objects']['disk']['radius'] = sc['objects']['disk']['radius'] * scale
    for chunk, mode in ((0, 0), (64, 0), (0, 4), (64, 4)):
        L.surf_set_kernel_timing(1)
        with torch.no_grad():
            for _ in range(5):
                r = surf_renderer_b200.render(sc, _chunk_prims=chunk, _math_mode=mode)
        torch.cuda.synchronize()
        ms = L.surf_mean_kernel_ms(0, None)
        L.surf_set_kernel_timing(0)
        hit = float((r['depth'] <= 1000).float().mean())
        print('radius x%-6g chunk %4d mode %d  k_intersect %.4f ms  hit fraction %.3f' % (scale, chunk, mode, ms, hit), flush=True)
