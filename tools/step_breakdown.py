"""Where does a fwd+bwd step spend its time outside k_intersect?  CUDA events between phases + CPU timers."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import scene_io, surf_renderer_b200
from surf_renderer_b200 import scenes as synth
from surf_renderer_b200._lib import lib

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
scene = synth.config_e(m=100000, width=size, height=size)
dev = torch.device('cuda', 0)
sc = scene_io.clone_scene(scene, device=dev)
leaves = [sc['objects']['disk']['pos'], sc['objects']['disk']['normal'], sc['materials']['albedo'], sc['lights']['pos']]
for t in leaves: t.requires_grad_(True)
opt = torch.optim.Adam(leaves, lr=1e-4)
with torch.no_grad():
    tgt = surf_renderer_b200.render(scene_io.clone_scene(synth.config_e_target_scene(scene), device=dev))['image']
lib().surf_set_kernel_timing(1)
E = lambda: torch.cuda.Event(enable_timing=True)
def step(rec=None, cpu=None):
    t0 = time.perf_counter()
    if rec: rec[0].record()
    opt.zero_grad(set_to_none=True)
    res = surf_renderer_b200.render(sc, _math_mode=mode)
    t1 = time.perf_counter()
    if rec: rec[1].record()
    loss = ((res['image'] - tgt) ** 2).mean()
    loss.backward()
    t2 = time.perf_counter()
    if rec: rec[2].record()
    opt.step()
    t3 = time.perf_counter()
    if rec: rec[3].record()
    if cpu is not None: cpu.append((t1 - t0, t2 - t1, t3 - t2))
for _ in range(5): step()
torch.cuda.synchronize()
for rep in range(3):
    rec = [E() for _ in range(4)]; cpu = []
    torch.cuda.synchronize()
    step(rec, cpu)
    torch.cuda.synchronize()
    k = [lib().surf_last_kernel_ms(i) for i in range(3)]
    print('GPU ms: render %.3f (k_intersect %.3f, k_shade %.3f)  loss+backward %.3f (k_backward %.3f)  adam %.3f | total %.3f' % (
        rec[0].elapsed_time(rec[1]), k[0], k[1], rec[1].elapsed_time(rec[2]), k[2], rec[2].elapsed_time(rec[3]), rec[0].elapsed_time(rec[3])))
    print('CPU ms: render %.3f  loss+backward %.3f  adam %.3f' % tuple(1e3 * x for x in cpu[0]))
