"""Where the end-to-end (host buffers) step of config E spends its time: host wall clock of surf_step_host_begin (enqueue) and
surf_step_host_end (wait + D2H), and the library's CUDA-event timers of the three kernel stages inside the same steps."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                    # noqa: E402
import surf_renderer_b200                                       # noqa: E402
from surf_renderer_b200 import scenes as synth                  # noqa: E402
from surf_renderer_b200._lib import check, lib                  # noqa: E402
from surf_renderer_b200.marshal import Marshalled, make_options  # noqa: E402
from surf_renderer_b200.scenes import clone_scene               # noqa: E402

scene = synth.config_e()
with torch.no_grad():
    tgt = surf_renderer_b200.render(clone_scene(synth.config_e_target_scene(scene), device='cuda'))['image']
m = Marshalled(clone_scene(scene), 'cpu')
m.floats = [t.pin_memory() for t in m.floats]
m.ints = {k: v.pin_memory() for k, v in m.ints.items()}
m.cam_vecs = {k: v.pin_memory() for k, v in m.cam_vecs.items()}
n = 1024 * 1024
target = tgt.detach().reshape(-1, 3).cpu().pin_memory()
grads = [torch.zeros_like(t).pin_memory() for t in m.floats]
csc, ccam, copt, csg = m.c_scene(), m.c_camera(), make_options({}, (0, n)), m.c_grads(grads)
ctx = lib().surf_context_create(0)
loss = C.c_float()
for _ in range(3):
    check(lib().surf_step_host_begin(ctx, C.byref(csc), C.byref(ccam), C.byref(copt), target.data_ptr(), 1.0 / (3.0 * n)))
    check(lib().surf_step_host_end(ctx, C.byref(csg), C.byref(loss)))
lib().surf_set_kernel_timing(1)
tb = te = 0.0
N = 10
t0 = time.perf_counter()
for _ in range(N):
    a = time.perf_counter()
    check(lib().surf_step_host_begin(ctx, C.byref(csc), C.byref(ccam), C.byref(copt), target.data_ptr(), 1.0 / (3.0 * n)))
    b = time.perf_counter()
    check(lib().surf_step_host_end(ctx, C.byref(csg), C.byref(loss)))
    c = time.perf_counter()
    tb += b - a
    te += c - b
total = (time.perf_counter() - t0) / N
print('e2e step %.3f ms: host_begin returns after %.3f ms, host_end takes %.3f ms; stages: intersect %.3f shade %.4f backward %.4f ms' % (
    total * 1e3, tb / N * 1e3, te / N * 1e3, lib().surf_mean_kernel_ms(0, None), lib().surf_mean_kernel_ms(1, None), lib().surf_mean_kernel_ms(2, None)))
