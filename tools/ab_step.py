"""A/B timing of the per-pixel kernels inside the fused step (config E, or config D with `d`): prints the mean
k_intersect / k_shade / k_backward durations (CUDA events inside the library) for the build SURF_B200_LIB points at.
  SURF_B200_LIB=build/variants/lib_x.so python tools/ab_step.py [e|d] [reps]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                      # noqa: E402
import surf_renderer_b200         # noqa: E402
from surf_renderer_b200 import dist as sdist, scenes as synth       # noqa: E402
from surf_renderer_b200._lib import lib, LIB_PATH                   # noqa: E402
from surf_renderer_b200.scenes import clone_scene                   # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'e'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device('cuda', 0)
flush = torch.empty(144 << 20, dtype=torch.uint8, device=dev)
if which == 'e':
    scene = synth.config_e()
    tgt = surf_renderer_b200.render(clone_scene(synth.config_e_target_scene(scene), device=dev))['image'].detach()
    sc = clone_scene(scene, device=dev)
    for t in (sc['objects']['disk']['pos'], sc['objects']['disk']['normal'], sc['materials']['albedo'], sc['lights']['pos']):
        t.requires_grad_(True)
    extra = {}
    if os.environ.get('AB_PPT'):
        extra['_pixels_per_thread'] = int(os.environ['AB_PPT'])
    if os.environ.get('AB_CHUNK'):
        extra['_chunk_prims'] = int(os.environ['AB_CHUNK'])
    plan = surf_renderer_b200.MSEStep(sc, tgt, **extra)
    step = plan
else:
    host = synth.config_d_batch(64)
    plan = sdist.ShardedBatchStep(host, device=dev, double_sided=True)
    w = torch.rand(64, 128, 128, 3, device=dev)
    step = lambda: plan.step(lambda im: (im * w).sum())      # noqa: E731
for _ in range(3):
    step()
torch.cuda.synchronize()
lib().surf_set_kernel_timing(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    flush.fill_(i & 255)
    step()
e1.record()
torch.cuda.synchronize()
n = C.c_int32()
out = {'lib': os.path.basename(LIB_PATH), 'workload': which, 'ppt': os.environ.get('AB_PPT'), 'chunk': os.environ.get('AB_CHUNK'), 'ms_per_step': e0.elapsed_time(e1) / reps,
       'intersect_ms': lib().surf_mean_kernel_ms(0, C.byref(n)), 'shade_ms': lib().surf_mean_kernel_ms(1, C.byref(n)),
       'backward_ms': lib().surf_mean_kernel_ms(2, C.byref(n)), 'launches_timed': n.value}
print(json.dumps(out))
